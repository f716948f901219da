"""The reference's OWN calls into its absent natives, replayed through the product's `vacmap_index` shim.

tests/golden/native_calls.json.gz holds every call the reference's get_readmap_DP_test made into `index.map`,
`edlib.align` and `mp.k_cigar` for the first reads of each bulk case (recorded in the build container while the
reference ran over the oracle natives, tests/golden/make_bulk.py), with the result it got.  Here each call goes through
vacmap_b200.vacmap_index (libvacmap_b200.so, CUDA) and must return the same thing.  Call by call equal natives mean the
reference's Python over the CUDA natives is, by induction over its call sequence, the run the fixtures were made from
(the reference itself cannot be imported on the GPU box)."""
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

import bulk

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
CALLS = json.load(gzip.open(os.path.join(HERE, "golden", "native_calls.json.gz"), "rt"))["cases"]


@pytest.mark.parametrize("name", sorted(CALLS))
def test_recorded_native_calls_replay_identically(gpu_ctx, name):
    from vacmap_b200 import vacmap_index as vi
    mode, k, w, over = bulk.CASES[name]
    ref = bulk.reference_for(name)
    reads = bulk.reads_for(name, ref)
    A = vi.Aligner(contigs=ref, w=w, k=k)
    assert A.k == k and [x[0].decode() for x in A.seq_offset] == [n for n, _ in ref]
    n_map = n_ed = n_ext = n_fill = 0
    for row in CALLS[name]:
        seq = reads[row["i"]][1].upper()
        ext, fill, ed = [], [], []
        for c in row["calls"]:
            if c["f"] == "map":
                got = A.map(seq, check_num=c["check_num"], mid_occ=c["mid_occ"])
                arr = np.array(got, dtype=np.int64).reshape(-1, 4)
                assert len(got) == c["n"] and hashlib.sha1(arr.tobytes()).hexdigest()[:16] == c["sha"], (name, row["i"])
                n_map += 1
            elif c["f"] == "edlib":
                ed.append(c)
            elif c["kw"].get("bw") == 100:
                ext.append(c)
            else:
                fill.append(c)
        # one launch per kind and read (the shim's *_batch forms); a few calls also one by one
        if ed:
            d = vi.edlib_align_batch([(c["q"], c["t"]) for c in ed])
            assert [x["editDistance"] for x in d] == [c["out"] for c in ed]
            assert vi.edlib_align(query=ed[0]["q"], target=ed[0]["t"], task="distance")["editDistance"] == ed[0]["out"]
            n_ed += len(ed)
        if ext:
            kw = dict(ext[0]["kw"])
            r = vi.k_cigar_batch([(c["t"], c["q"]) for c in ext], **kw)
            assert [(x[2], x[3]) for x in r] == [(c["out"][2], c["out"][3]) for c in ext]     # q_e, t_e: all the path reads
            one = vi.k_cigar(ext[0]["t"], ext[0]["q"], **kw)
            assert (one[2], one[3]) == (ext[0]["out"][2], ext[0]["out"][3])
            n_ext += len(ext)
        if fill:
            kw = dict(fill[0]["kw"])
            r = vi.k_cigar_batch([(c["t"], c["q"]) for c in fill], **kw)
            assert [list(x) for x in r] == [c["out"] for c in fill]
            assert list(vi.k_cigar(fill[-1]["t"], fill[-1]["q"], **kw)) == fill[-1]["out"]
            n_fill += len(fill)
    assert n_map >= 10 and n_ed >= 10 and n_fill >= 50 and n_ext + n_fill >= 100
