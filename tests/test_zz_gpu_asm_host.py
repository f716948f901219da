"""asm mode: the product host loop over the CUDA linked DPs.  Kept in a file that sorts last: it was added after the
round's GPU budget was spent (its pieces -- the kernels through the oracle's loop, the host loop with the oracle's DP
-- were each run), so a surprise here must not hide the rest of the GPU suite behind `-x`."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_asm_host_loop_over_the_cuda_linked_dps(gpu_ctx):
    """vacmap_b200.asm.linked_chain_path (product: batch loop + carry + traceback over vm_chain_linked_batch) gives
    the reference-pinned paths of both rounds, bail-out flow and traceback quirk included."""
    import os
    from vacmap_b200 import asm
    from vacmap_b200.chain import ChainParams
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "asm_linked.npz"))
    p1 = ChainParams(kmersize=15, skipcost=40.0, maxdiff=50, maxgap=1000)
    for fi in range(int(G["n_flows"])):
        batches = [G["f%d_b%d" % (fi, bi)].astype(np.int64).reshape(-1, 4) for bi in range(int(G["f%d_nb" % fi]))]
        path = asm.linked_chain_path(batches, p1, ctx=gpu_ctx)
        assert np.array_equal(np.array(path, dtype=np.int64).reshape(-1, 4), G["f%d_path" % fi]), fi
    p2 = ChainParams(kmersize=9, skipcost=30.0, maxdiff=30, maxgap=99)
    for fi in range(int(G["n_lflows"])):
        batches = [G["l%d_b%d" % (fi, bi)].astype(np.int64).reshape(-1, 4) for bi in range(int(G["l%d_nb" % fi]))]
        if int(G["l%d_err" % fi]):
            with pytest.raises(IndexError):
                asm.linked_chain_path(batches, p2, second_round=True, ctx=gpu_ctx)
            continue
        path = asm.trim_overlaps(asm.linked_chain_path(batches, p2, second_round=True, ctx=gpu_ctx))
        assert np.array_equal(np.array(path, dtype=np.int64).reshape(-1, 4), G["l%d_path" % fi]), fi


def test_asm_product_path_matches_reference_end_to_end(gpu_ctx):
    """`vacmap_b200.asm.assembly_align` -- seeding batches, both linked-DP rounds, the re-seeding between them,
    ass_extend_func with asm's own rebuild / inversion fix / split + link_cigar, record assembly, all over the CUDA entry
    points -- gives the rows and CIGARs the REFERENCE's assembly_get_readmap_DP_test gave for the 520 kb contig read with an
    inversion, a deletion and an insertion (tests/golden/asm_e2e.json.gz), with and without --eqx; and the SAM text of
    the mode's emitter on them equals the reference's lines (tests/golden/asm_sam.json.gz)."""
    import gzip
    import json
    import os
    import synth
    import vacmap_b200 as vb
    from vacmap_b200 import asm, sam
    here = os.path.dirname(os.path.abspath(__file__))
    E = json.load(gzip.open(os.path.join(here, "golden", "asm_e2e.json.gz"), "rt"))
    ref, read = synth.asm_e2e_inputs()
    ix = vb.Index(ref, w=10, k=15, ctx=gpu_ctx)
    for case in E["cases"]:
        opt = vb.default_option("S", eqx=case["eqx"])
        opt.update({"golbal_skipcost": 30., "golbal_maxdiff": 50, "local_skipcost": 30., "local_maxdiff": 30, "local_kmersize": 9})
        got = asm.assembly_align("ctgread", read, ix, opt, ctx=gpu_ctx)
        assert [list(r) for r in got] == case["records"], case["eqx"]
    # a contig below 500 kb takes the per-read path (mammap_asm.py:23205-23207)
    short = asm.assembly_align("short", read[:60000], ix, vb.default_option("S"), ctx=gpu_ctx)
    assert len(short) >= 1 and short[0][1] == "chr1"
    ix.close()
