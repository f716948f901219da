"""GPU parity on the bulk fixture: all 2 136 reads of tests/golden/bulk_e2e.json.gz -- records produced by the
REFERENCE's own get_readmap_DP_test (mode H, H --eqx --MD, L at -k 19 -w 10, S on nested-SV reads; reads it leaves
unmapped; up to 26 records per read) -- against the CUDA pipeline through the C ABI: coordinates, strand, MAPQ, CIGAR
(length + sha1), the SAM text (sha1 of the reference's lines), the per-read status, and the BRANCH COUNTERS: the
library must have gone through the second extension pass, drop_misplaced, fix_simple_inv, merge_conjacent and the
heuristic global DP exactly as often as the reference did on the same reads."""
import hashlib

import numpy as np
import pytest

import bulk
import refrun_options
from test_oracle_bulk import BULK

pytestmark = pytest.mark.gpu

UNMAPPED = (1, 2, 3, 5)      # VM_READ_FEW_ANCHORS / LOW_SCORE / SHORT_LOCAL / NO_RECORDS


@pytest.mark.parametrize("name", sorted(BULK["cases"]))
def test_cuda_matches_reference_on_all_bulk_reads(gpu_ctx, name):
    import vacmap_b200 as vb
    from vacmap_b200 import sam
    case = BULK["cases"][name]
    ref = bulk.reference_for(name)
    reads = bulk.reads_for(name, ref)
    opt = refrun_options.default_option(case["mode"], **case["opt"])
    ix = vb.Index(ref, w=case["w"], k=case["k"], ctx=gpu_ctx)
    al = vb.Aligner(ix, opt, case["mode"])
    got = al.align_batch(reads)
    status = al.last_status
    counters = dict(al.last_stage_ms)
    contig2seq = {n: ix.seq(n) for n, _ in ref}
    contig2iloc = {n: i for i, (n, _) in enumerate(ref)}
    bad = []
    for i, ((rid, seq), g, want) in enumerate(zip(reads, got, case["reads"])):
        if [bulk.squash(r) for r in g] != want["records"]:
            bad.append(rid)
            continue
        if want["status"] == "ok":
            assert status[i] == 0, rid
            lines = sam.get_bam_dict_str(g, seq.upper(), None, contig2iloc, contig2seq, opt["md"], opt["shortcs"], opt["cigar2cg"],
                                         opt["markunbalancetra"], opt)
            assert hashlib.sha1("\n".join(lines).encode()).hexdigest()[:16] == want["sam"], rid
        elif want["status"] == "unmapped":
            assert status[i] in UNMAPPED, (rid, status[i])
        else:
            assert status[i] == 4, (rid, status[i], want["status"])      # the reference raised inside the read
    assert not bad, "%d of %d reads differ from the reference: %s" % (len(bad), len(reads), bad[:10])
    tot = case["totals"]
    for key, ckey in (("second_pass", "c_second_pass"), ("drop_misplaced", "c_drop_misplaced"), ("fix_simple_inv", "c_fix_simple_inv"),
                      ("merge_conjacent", "c_merge_conjacent"), ("fast_global", "c_fast_global")):
        assert int(counters.get(ckey, 0)) == tot.get(key, 0), (key, counters.get(ckey), tot.get(key))
    # the multi-chain `_mismatch` local DP: taken for reads that still have > 1 guide chain after merge_chain /
    # drop_somechains (inside the reference's njit function, where it cannot be counted): bounded by the reads
    # that entered with > 1 chain
    assert 0 < int(counters.get("c_mismatch_dp", 0)) <= tot["n_chains"]
    ix.close()


def test_lockstep_and_pipelined_agree_on_bulk(gpu_ctx):
    """The worker-pool path and the lock-step path give the same records and the same branch counts."""
    import vacmap_b200 as vb
    name = "bulk_S"
    case = BULK["cases"][name]
    ref = bulk.reference_for(name)
    reads = bulk.reads_for(name, ref)[:160]
    opt = refrun_options.default_option(case["mode"], **case["opt"])
    ix = vb.Index(ref, w=case["w"], k=case["k"], ctx=gpu_ctx)
    a = vb.Aligner(ix, opt, case["mode"], workers=1)
    ra = a.align_batch(reads)
    b = vb.Aligner(ix, opt, case["mode"], workers=4, chunk_reads=37)
    rb = b.align_batch(reads)
    assert ra == rb and (a.last_status == b.last_status).all()
    for k in ("c_merge_conjacent", "c_fix_simple_inv", "c_mismatch_dp"):
        assert a.last_stage_ms.get(k, 0) == b.last_stage_ms.get(k, 0), k
    ix.close()
