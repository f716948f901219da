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
