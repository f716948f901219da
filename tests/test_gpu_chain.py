"""GPU parity: CUDA global chaining (through the C ABI) vs the oracle and the golden vectors."""
import os

import numpy as np
import pytest

import oracle
import synth

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "chain.npz"))


def _check_against_oracle(anchor_list, read_lens, results, params):
    nfast = 0
    for a, L, r in zip(anchor_list, read_lens, results):
        a = np.asarray(a, dtype=np.int64)
        perm = oracle.argsort_i64(a[:, 0])
        srt = a[perm]
        assert (r.sorted == srt).all()
        fast = len(a) / L > 5
        if not fast:
            g, S, P, A, op = oracle.chain_global_d_all(srt, params.kmersize, params.skipcost, params.maxdiff, params.maxgap)
            if g == -1:
                fast = True
        if fast:
            g, S, P, A = oracle.chain_fast(srt, params.kmersize, 0, params.skipcost, params.maxdiff, params.maxgap)
            nfast += 1
        assert r.used_fast == fast
        assert r.g_max_index == g
        assert (r.S == S).all(), np.flatnonzero(r.S != S)[:5]
        assert (r.P == P).all()
        assert (r.S_arg == A).all()
    return nfast


def test_golden_global(gpu_ctx):
    import vacmap_b200 as vb
    prm = vb.ChainParams()
    al, Ls = [], []
    for ci in range(int(G["g_count"])):
        a = G["g_%d_raw" % ci].astype(np.int64)
        al.append(a)
        Ls.append(15000 if len(a) <= 5 * 15000 else 100)
    res = vb.chain_global_batch(al, Ls, prm, ctx=gpu_ctx)
    for ci, r in enumerate(res):
        ge = int(G["g_%d_exact" % ci])
        assert (r.sorted == G["g_%d_a" % ci]).all()
        if ge >= 0:
            assert not r.used_fast and r.g_max_index == ge
            assert (r.S == G["g_%d_S" % ci]).all() and (r.P == G["g_%d_P" % ci]).all() and (r.S_arg == G["g_%d_A" % ci]).all()
        else:
            assert r.used_fast and r.g_max_index == int(G["g_%d_fg" % ci])
            assert (r.S == G["g_%d_fS" % ci]).all() and (r.P == G["g_%d_fP" % ci]).all() and (r.S_arg == G["g_%d_fA" % ci]).all()
    # the same anchors forced down the fast path (n / read_len > 5)
    res = vb.chain_global_batch(al, [max(1, len(a) // 6) for a in al], prm, ctx=gpu_ctx)
    for ci, r in enumerate(res):
        assert r.used_fast and r.g_max_index == int(G["g_%d_fg" % ci])
        assert (r.S == G["g_%d_fS" % ci]).all() and (r.P == G["g_%d_fP" % ci]).all() and (r.S_arg == G["g_%d_fA" % ci]).all()


def test_random_batch_vs_oracle(gpu_ctx):
    import vacmap_b200 as vb
    rng = np.random.default_rng(7)
    prm = vb.ChainParams()
    al, Ls = [], []
    for t in range(160):
        if t % 5 == 0:
            a = synth.anchors_tieheavy(rng, n=int(rng.integers(3, 700)))
            L = 2000
        else:
            a = synth.anchors_global(rng, n_true=int(rng.integers(5, 700)), n_noise=int(rng.integers(0, 2500)))
            L = 15000
        al.append(a)
        Ls.append(L)
    # edge cases: tiny reads, equal positions only, large (beyond the smem classes)
    al.append(np.array([[5, 100, 1, 15]], dtype=np.int64)); Ls.append(100)
    al.append(np.array([[5, 100, 1, 15], [5, 300, -1, 15], [5, 100, 1, 15]], dtype=np.int64)); Ls.append(100)
    al.append(synth.anchors_global(rng, n_true=9000, n_noise=9000, L=200000)); Ls.append(200000)
    al.append(synth.anchors_global(rng, n_true=0, n_noise=2600, repeats=False)); Ls.append(15000)
    res = vb.chain_global_batch(al, Ls, prm, ctx=gpu_ctx)
    nfast = _check_against_oracle(al, Ls, res, prm)
    assert nfast >= 1


def test_empty_and_ragged(gpu_ctx):
    import vacmap_b200 as vb
    rng = np.random.default_rng(3)
    al = [np.zeros((0, 4), np.int64), synth.anchors_global(rng, n_true=50, n_noise=10), np.zeros((0, 4), np.int64)]
    res = vb.chain_global_batch(al, [100, 15000, 100], ctx=gpu_ctx)
    assert len(res[0].S) == 0 and len(res[2].S) == 0
    _check_against_oracle(al[1:2], [15000], res[1:2], vb.ChainParams())
    assert vb.chain_global_batch([], [], ctx=gpu_ctx) == []


def test_full_size_properties(gpu_ctx):
    """BASELINE config[1] scale (10k reads): size-independent properties of the result."""
    import vacmap_b200 as vb
    rng = np.random.default_rng(11)
    base = [synth.anchors_global(rng, n_true=int(rng.integers(300, 700)), n_noise=int(rng.integers(100, 1500)))
            for _ in range(200)]
    al = [base[i % len(base)] for i in range(10000)]
    res = vb.chain_global_batch(al, [15000] * len(al), ctx=gpu_ctx)
    for i, r in enumerate(res):
        n = len(r.S)
        # sortedness of the replayed argsort and of S_arg; P points backwards; g_max is the first maximum
        assert (np.diff(r.sorted[:, 0]) >= 0).all()
        assert sorted(r.S_arg.tolist()) == list(range(n))
        assert (np.diff(r.S[r.S_arg]) >= 0).all()
        ok = (r.P == vb._lib.NOPRE) | ((r.P >= 0) & (r.P < np.arange(n)))
        assert ok.all()
        assert r.g_max_index == int(np.argmax(r.S))
        # identical inputs give identical outputs (idempotence across the batch)
        j = i % len(base)
        if i >= len(base):
            assert (r.S == res[j].S).all() and (r.P == res[j].P).all() and (r.S_arg == res[j].S_arg).all()


def test_golden_local(gpu_ctx):
    """Local DPs through vm_chain_local_batch vs what the REFERENCE's own numba functions returned for the same sorted
    anchors (tests/golden/chain.npz): _fine_list / _fine_list_mismatch (modes H and L parameters), exact score and the
    trimmed path; the _fast twins' score; and, on unsorted input, the oracle after the replayed argsort."""
    import vacmap_b200 as vb
    n = int(G["l_count"])
    al = [G["l_%d_a" % ci].astype(np.int64) for ci in range(n)]
    Ls = [int((a[:, 0] + a[:, 3]).max()) + 1 for a in al]
    for tag, var, sk, mg in (("fl", 1, 40.0, 99), ("flm", 2, 40.0, 99), ("fl59", 1, 59.0, 50)):
        prm = vb.ChainParams(kmersize=9, skipcost=sk, maxdiff=30, maxgap=mg, variant=var)
        res = vb.chain_local_batch(al, Ls, prm, presorted=True, ctx=gpu_ctx)
        for ci, r in enumerate(res):
            assert r.score == float(G["l_%d_%s_score" % (ci, tag)]), (tag, ci)
            want = G["l_%d_%s_path" % (ci, tag)].astype(np.int64).reshape(-1, 4)[::-1]      # the reference lists it descending
            assert r.path.shape == want.shape and (r.path == want).all(), (tag, ci)
            assert not r.used_fast
    for tag, var in (("flf", 1), ("flmf", 2)):
        prm = vb.ChainParams(kmersize=9, skipcost=40.0, maxdiff=30, maxgap=99, variant=var)
        res = vb.chain_local_batch(al, Ls, prm, presorted=True, force_fast=True, ctx=gpu_ctx)
        for ci, r in enumerate(res):
            assert r.used_fast and r.score == float(G["l_%d_%s_score" % (ci, tag)]), (tag, ci)
            want = G["l_%d_%s_path" % (ci, tag)].astype(np.int64).reshape(-1, 4)[::-1]
            assert r.path.shape == want.shape and (r.path == want).all(), (tag, ci)
    # unsorted input: the library's own argsort (numba quicksort replay on x + len), then the oracle on that order
    rng = np.random.default_rng(12)
    raw = [a[rng.permutation(len(a))] for a in al] + [np.zeros((0, 4), np.int64)]
    prm = vb.ChainParams(kmersize=9, skipcost=40.0, maxdiff=30, maxgap=99, variant=1)
    res = vb.chain_local_batch(raw, Ls + [100], prm, presorted=False, ctx=gpu_ctx)
    for a, r in zip(raw, res):
        if len(a) == 0:
            assert len(r.path) == 0
            continue
        srt = a[oracle.argsort_i64(a[:, 0] + a[:, 3])]
        sc, path, S, P, used_fast = oracle.chain_local(srt, 9, 1, 40.0, 30, 99)
        assert r.score == sc and (r.path == np.asarray(path, np.int64).reshape(-1, 4)[::-1]).all()


def test_asm_linked_dp_matches_reference_and_oracle(gpu_ctx):
    """asm mode: vm_chain_linked_batch == the reference's linked_..._d_all on the golden batch flows
    (tests/golden/asm_linked.npz: S float64 exact, P, S_arg, g_max_index, carried prefixes included), driven through
    the oracle's transcription of the batch loop; then all the recorded calls once more as ONE batch of jobs."""
    import os
    import oracle
    import oracle.asm as oasm
    from vacmap_b200.chain import ChainParams, chain_linked_batch
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "asm_linked.npz"))
    prm = ChainParams(kmersize=15, skipcost=40.0, maxdiff=50, maxgap=1000)
    jobs, want, n_fast = [], [], []
    for fi in range(int(G["n_flows"])):
        batches = [G["f%d_b%d" % (fi, bi)].astype(np.int64).reshape(-1, 4) for bi in range(int(G["f%d_nb" % fi]))]
        seen = []

        def dp(gs, gi, pS, pP, prl, lk, seen=seen, fi=fi):
            ci = len(seen)
            r = chain_linked_batch([(gs, gi, pS, pP, prl, lk)], prm, ctx=gpu_ctx)[0]
            # where the reference's exact DP bails out on opcount its caller takes the heuristic twin: so does the entry point
            tag = "f" if int(G["f%d_c%d_g" % (fi, ci)]) == -1 else ""
            assert r.used_fast == (1 if tag else 0), (fi, ci)
            n_fast.append(r.used_fast)
            assert r.g_max_index == int(G["f%d_c%d_%sg" % (fi, ci, tag)]), (fi, ci)
            assert np.array_equal(r.S, G["f%d_c%d_%sS" % (fi, ci, tag)]), (fi, ci)
            assert np.array_equal(r.P, G["f%d_c%d_%sP" % (fi, ci, tag)]), (fi, ci)
            assert np.array_equal(r.S_arg, G["f%d_c%d_%sA" % (fi, ci, tag)]), (fi, ci)
            jobs.append((gs, gi, pS.copy(), pP.copy(), prl, lk.copy()))
            want.append(r)
            seen.append(1)
            return r.g_max_index, r.S, r.P, r.S_arg

        path = oasm.first_round_path(batches, 15, 40., 50, 1000, dp=dp)
        assert np.array_equal(np.array(path, dtype=np.int64).reshape(-1, 4), G["f%d_path" % fi]), fi
    assert len(jobs) >= 20 and sum(n_fast) >= 1
    for r, w, j in zip(chain_linked_batch(jobs, prm, ctx=gpu_ctx), want, jobs):
        assert r.g_max_index == w.g_max_index and np.array_equal(r.S, w.S) and np.array_equal(r.P, w.P) and np.array_equal(r.S_arg, w.S_arg)
        if r.used_fast:
            o = oracle.chain_linked_fast(j[0], j[1], j[2], j[3], j[4], j[5], 15, 40., 50, 1000)
        else:
            o = oracle.chain_linked_d_all(j[0], j[1], j[2], j[3], j[4], j[5], 15, 40., 50, 1000)
        assert o[0] == r.g_max_index and np.array_equal(o[1], r.S) and np.array_equal(o[2], r.P) and np.array_equal(o[3], r.S_arg)


def test_asm_linked_second_round_dp_matches_reference(gpu_ctx):
    """vm_chain_linked_batch with variant 4 == the reference's linked_..._fine_list_all on the golden second-round
    flows (local anchors, asm's read-gap table, carried prefixes), call by call through the oracle's loop."""
    import os
    import oracle.asm as oasm
    from vacmap_b200.chain import ChainParams, chain_linked_batch
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "asm_linked.npz"))
    prm = ChainParams(kmersize=9, skipcost=30.0, maxdiff=30, maxgap=99, variant=4)
    n_calls = 0
    for fi in range(int(G["n_lflows"])):
        batches = [G["l%d_b%d" % (fi, bi)].astype(np.int64).reshape(-1, 4) for bi in range(int(G["l%d_nb" % fi]))]
        seen = []

        def dp(gs, gi, pS, pP, prl, lk, seen=seen, fi=fi):
            ci = len(seen)
            r = chain_linked_batch([(gs, gi, pS, pP, prl, lk)], prm, ctx=gpu_ctx)[0]
            assert r.used_fast == 0 and r.g_max_index == int(G["l%d_c%d_g" % (fi, ci)]), (fi, ci)
            assert np.array_equal(r.S, G["l%d_c%d_S" % (fi, ci)]), (fi, ci)
            assert np.array_equal(r.P, G["l%d_c%d_P" % (fi, ci)]), (fi, ci)
            assert np.array_equal(r.S_arg, G["l%d_c%d_A" % (fi, ci)]), (fi, ci)
            seen.append(1)
            return r.g_max_index, r.S, r.P, r.S_arg

        if int(G["l%d_err" % fi]):
            with pytest.raises(IndexError):
                oasm.second_round_path(batches, 9, 30., 30, 99, dp=dp)
        else:
            path = oasm.second_round_path(batches, 9, 30., 30, 99, dp=dp)
            assert np.array_equal(np.array(path, dtype=np.int64).reshape(-1, 4), G["l%d_path" % fi]), fi
        n_calls += len(seen)
    assert n_calls >= 10
