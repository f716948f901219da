"""Oracle (C restatement) vs golden vectors produced by the reference's own numba code."""
import os

import numpy as np
import pytest

import oracle

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "chain.npz"))


def test_argsort_matches_numba_quicksort():
    for i in range(int(G["sort_count"])):
        a = G["sort_%d_in" % i]
        assert (oracle.argsort_i64(a) == G["sort_%d_out" % i]).all(), i
        assert (oracle.argsort_f64(a.astype(np.float64) / 7) == G["sort_%d_out" % i]).all(), i


def test_tables_shape():
    t = oracle.tables()
    assert len(t["extra"]) == 162756 and t["extra"].dtype == np.float32 and t["extra"][-1] == 36
    assert len(t["readgapcost"]) == 100 and t["readgapcost"][0] == 0
    assert len(t["log2cache"]) == 100000


@pytest.mark.parametrize("ci", range(int(G["g_count"])))
def test_global_exact_and_fast(ci):
    a = G["g_%d_a" % ci].astype(np.int64)
    g, S, P, A, op = oracle.chain_global_d_all(a, 15, 40.0, 50, 1000)
    assert g == int(G["g_%d_exact" % ci])
    if g >= 0:
        assert (S == G["g_%d_S" % ci]).all() and (P == G["g_%d_P" % ci]).all() and (A == G["g_%d_A" % ci]).all()
    g, S, P, A = oracle.chain_fast(a, 15, 0, 40.0, 50, 1000)
    assert g == int(G["g_%d_fg" % ci])
    assert (S == G["g_%d_fS" % ci]).all() and (P == G["g_%d_fP" % ci]).all() and (A == G["g_%d_fA" % ci]).all()


def test_global_bailout_case_present():
    assert any(int(G["g_%d_exact" % ci]) == -1 for ci in range(int(G["g_count"])))


@pytest.mark.parametrize("ci", range(int(G["l_count"])))
def test_local_variants(ci):
    a = G["l_%d_a" % ci].astype(np.int64)
    for tag, var, sk, mg in (("fl", 1, 40.0, 99), ("flm", 2, 40.0, 99), ("fl59", 1, 59.0, 50)):
        sc, path, S, P, used_fast = oracle.chain_local(a, 9, var, sk, 30, mg)
        assert sc == float(G["l_%d_%s_score" % (ci, tag)])
        assert (path == G["l_%d_%s_path" % (ci, tag)]).all()
    for tag, var in (("flf", 1), ("flmf", 2)):
        g, S, P, A = oracle.chain_fast(a, 9, var, 40.0, 30, 99)
        assert S[g] == float(G["l_%d_%s_score" % (ci, tag)])


def test_live_reference_if_present():
    """When the reference tree is importable (build container), compare on fresh random inputs."""
    import refimport
    if not refimport.available():
        pytest.skip("reference tree not present")
    import synth
    m = refimport.load_mode("clrnano")
    f = m.get_optimal_chain_sortbyreadpos_forSV_inv_test_merged_fine_list_d_all
    rng = np.random.default_rng(99)
    for t in range(4):
        a = synth.anchors_global(rng, n_true=200, n_noise=300)
        a = a[oracle.argsort_i64(a[:, 0])]
        g, S, P, A, _ = f(a, kmersize=15, skipcost=40., maxdiff=50, maxgap=1000)
        g2, S2, P2, A2, _ = oracle.chain_global_d_all(a, 15, 40.0, 50, 1000)
        assert g == g2 and (S == S2).all() and (P == P2).all() and (A == A2).all()
