"""asm mode, first brick: the oracle's linked global DP with carry-in and the batch loop around it
against the reference's own njit function (tests/golden/asm_linked.npz, made by tests/golden/make_golden.py asm)."""
import os

import numpy as np

import oracle
import oracle.asm as oasm

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "asm_linked.npz"))


def test_linked_dp_calls_match_reference_bit_for_bit():
    n_calls = n_carry = 0
    bailed = []
    for fi in range(int(G["n_flows"])):
        batches = [G["f%d_b%d" % (fi, bi)].astype(np.int64).reshape(-1, 4) for bi in range(int(G["f%d_nb" % fi]))]
        seen = []

        def dp(gs, gi, pS, pP, prl, lk, seen=seen, fi=fi):
            ci = len(seen)
            head = G["f%d_c%d_head" % (fi, ci)]
            assert (float(gs), float(gi), float(prl)) == tuple(head), (fi, ci)
            assert np.array_equal(pS, G["f%d_c%d_preS" % (fi, ci)]) and np.array_equal(pP, G["f%d_c%d_preP" % (fi, ci)]), (fi, ci)
            g, S, P, A, _ = oracle.chain_linked_d_all(gs, gi, pS, pP, prl, lk, 15, 40., 50, 1000)
            assert g == int(G["f%d_c%d_g" % (fi, ci)]), (fi, ci)
            if g >= 0:                      # after the opcount bail-out the reference's arrays are only partly written
                assert np.array_equal(S, G["f%d_c%d_S" % (fi, ci)]), (fi, ci)          # float64, exact
                assert np.array_equal(P, G["f%d_c%d_P" % (fi, ci)]), (fi, ci)
                assert np.array_equal(A, G["f%d_c%d_A" % (fi, ci)]), (fi, ci)
            # the heuristic twin on the same arguments (what the loop calls after a bail-out)
            fg, fS, fP, fA = oracle.chain_linked_fast(gs, gi, pS, pP, prl, lk, 15, 40., 50, 1000)
            assert fg == int(G["f%d_c%d_fg" % (fi, ci)]), (fi, ci)
            assert np.array_equal(fS, G["f%d_c%d_fS" % (fi, ci)]), (fi, ci)
            assert np.array_equal(fP, G["f%d_c%d_fP" % (fi, ci)]), (fi, ci)
            assert np.array_equal(fA, G["f%d_c%d_fA" % (fi, ci)]), (fi, ci)
            seen.append(len(pS))
            bailed.append(g == -1)
            return g, S, P, A

        path = oasm.first_round_path(batches, 15, 40., 50, 1000, dp=dp)
        assert len(seen) == int(G["f%d_calls" % fi])
        assert np.array_equal(np.array(path, dtype=np.int64).reshape(-1, 4), G["f%d_path" % fi]), fi
        n_calls += len(seen)
        n_carry += sum(1 for s in seen if s > 0)
    assert n_calls >= 20 and n_carry >= 10 and sum(bailed) >= 1


def test_default_dp_is_the_oracle_and_short_chains_give_nothing():
    batches = [G["f2_b%d" % bi].astype(np.int64) for bi in range(int(G["f2_nb"]))]
    assert np.array_equal(np.array(oasm.first_round_path(batches, 15, 40., 50, 1000)), G["f2_path"])
    assert oasm.first_round_path([np.array([[5, 100, 1, 15]])], 15, 40., 50, 1000) == []
    assert oasm.first_round_path([], 15, 40., 50, 1000) == []


def test_second_round_linked_dp_matches_reference():
    """`linked_..._fine_list_all` (local anchors, asm's own read-gap table, no bail-out) call by call, and the trimmed
    ascending path of the second-round loop (the trimming itself is the build's restatement of inline code)."""
    n_calls = n_err = 0
    for fi in range(int(G["n_lflows"])):
        batches = [G["l%d_b%d" % (fi, bi)].astype(np.int64).reshape(-1, 4) for bi in range(int(G["l%d_nb" % fi]))]
        seen = []

        def dp(gs, gi, pS, pP, prl, lk, seen=seen, fi=fi):
            ci = len(seen)
            g, S, P, A, _ = oracle.chain_linked_d_all(gs, gi, pS, pP, prl, lk, 9, 30., 30, 99, local=True)
            assert g == int(G["l%d_c%d_g" % (fi, ci)]), (fi, ci)
            assert np.array_equal(S, G["l%d_c%d_S" % (fi, ci)]), (fi, ci)
            assert np.array_equal(P, G["l%d_c%d_P" % (fi, ci)]), (fi, ci)
            assert np.array_equal(A, G["l%d_c%d_A" % (fi, ci)]), (fi, ci)
            seen.append(1)
            return g, S, P, A

        if int(G["l%d_err" % fi]):
            # a chain starting in a carried anchor: its negated "no predecessor" mark is followed as an index
            # (:23380-23385) and the reference raises -- so must the restatement
            import pytest
            with pytest.raises(IndexError):
                oasm.second_round_path(batches, 9, 30., 30, 99, dp=dp)
            n_err += 1
            n_calls += len(seen)
            continue
        path = oasm.second_round_path(batches, 9, 30., 30, 99, dp=dp)
        assert len(seen) == int(G["l%d_calls" % fi])
        want = G["l%d_path" % fi]
        assert np.array_equal(np.array(path, dtype=np.int64).reshape(-1, 4), want), fi
        if len(want) > 1:           # ascending, and no anchor reaches past its successor's start
            assert (np.diff(want[:, 0]) >= 0).all() and (want[:-1, 0] + want[:-1, 3] <= want[1:, 0]).all()
        n_calls += len(seen)
    assert n_calls >= 10


def test_second_round_reseeding_matches_reference():
    """`yield_second_mapinfo` / `collect_second_round_anchors` (mammap_asm.py:22444-22755) restated over the per-read
    oracle's window / scan pieces == the reference's own functions on a 60 kb contig read with an inversion and a
    deletion (tests/golden/asm_reseed.npz), for three batch sizes: same batches, same anchors, same order."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import oracle.pipeline as pl
    from vacmap_b200.sam import reverse_complement
    import synth
    R = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "asm_reseed.npz"))
    # the generator's seeded inputs (tests/golden/make_golden.py::asm_reseed_inputs), rebuilt here without the reference
    ref = synth.make_reference(77, 240000, n_contigs=2)
    rng = np.random.default_rng(78)
    src = np.frombuffer(ref[0][1].encode(), dtype=np.uint8)[20000:82000].copy()
    comp = np.zeros(256, np.uint8)
    for x, y in zip(b"ACGT", b"TGCA"):
        comp[x] = y
    parts = [src[:20000], comp[src[20000:26000]][::-1], src[26000:40000], src[43000:]]
    read = synth.mutate(rng, np.concatenate(parts), 0.01).tobytes().decode()
    # the first-round path the fixture was made from is itself reproduced by the oracle
    ox = oracle.Index(ref)
    a = np.array(ox.map(read, -1, -1), dtype=np.int64)
    a = a[oracle.argsort_i64(a[:, 0])]
    raw = np.array(oasm.first_round_path([a], 15, 40., 50, 1000)[::-1], dtype=np.int64)
    assert np.array_equal(raw, R["raw"])
    ctg = pl.Contigs([n for n, _ in ref], [s for _, s in ref])
    rc = reverse_complement(read)
    n_anchor = 0
    for bi in range(3):
        got = list(oasm.yield_second_mapinfo(raw, read, rc, ctg, 9, int(R["batch_%d" % bi])))
        assert len(got) == int(R["n_%d" % bi]), bi
        for ci, x in enumerate(got):
            assert np.array_equal(x, R["b%d_%d" % (bi, ci)]), (bi, ci)
            n_anchor += len(x)
    assert n_anchor > 10000
    # and the second-round chain over those batches covers the read on both strands (the inversion)
    path = oasm.second_round_path(got, 9, 30., 30, 99)
    strands = {p[2] for p in path}
    assert strands == {1, -1} and path[0][0] < 200 and path[-1][0] > len(read) - 300


def test_asm_end_to_end_matches_reference():
    """`oracle.asm.assembly_align` (both rounds, re-seeding, asm's own rebuild / inversion fix / split + link_cigar,
    record assembly) == the reference's assembly_get_readmap_DP_test on a 520 kb contig read with an inversion, a
    deletion and an insertion (tests/golden/asm_e2e.json.gz): same rows, same CIGARs, with and without --eqx."""
    import gzip
    import json
    import oracle.pipeline as pl
    import synth
    import test_oracle_e2e  # noqa: F401  (default options live next to the per-read fixtures)
    E = json.load(gzip.open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "asm_e2e.json.gz"), "rt"))
    ref, read = synth.asm_e2e_inputs()
    ox = oracle.Index(ref)
    ctg = pl.Contigs([n for n, _ in ref], [s for _, s in ref])
    for case in E["cases"]:
        opt = {"eqx": case["eqx"], "H": False, "golbal_skipcost": 30., "golbal_maxdiff": 50, "local_skipcost": 30.,
               "local_maxdiff": 30, "local_kmersize": 9}
        got = oasm.assembly_align("ctgread", read, ox, ctg, opt)
        assert [list(r) for r in got] == case["records"], case["eqx"]
    assert len(E["cases"][0]["records"]) == 3


def test_link_cigar_matches_reference():
    import json
    rows = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "asm_link_cigar.json")))
    assert len(rows) == 300 and any(r[2] != r[0] + r[1] for r in rows)
    for a, b, want in rows:
        assert oasm.link_cigar(a, b) == want, (a, b)


def test_asm_end_to_end_second_contig_matches_reference():
    """A second contig pinned by the reference's own run (tests/golden/asm_e2e2.json.gz): taken from the reverse strand,
    with a 30 kb piece of the other contig, a 6 kb tandem duplication and a 3 kb deletion -- records on two contigs and
    the reverse strand."""
    import gzip
    import json
    import oracle.pipeline as pl
    import synth
    E = json.load(gzip.open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "asm_e2e2.json.gz"), "rt"))
    ref, read = synth.asm_e2e_inputs_2()
    ox = oracle.Index(ref)
    ctg = pl.Contigs([n for n, _ in ref], [s for _, s in ref])
    for case in E["cases"]:
        opt = {"eqx": case["eqx"], "H": False, "golbal_skipcost": 30., "golbal_maxdiff": 50, "local_skipcost": 30.,
               "local_maxdiff": 30, "local_kmersize": 9}
        got = oasm.assembly_align("ctgread2", read, ox, ctg, opt)
        assert [list(r) for r in got] == case["records"], case["eqx"]
    recs = E["cases"][0]["records"]
    assert len(recs) == 5 and len({r[1] for r in recs}) == 2 and {r[2] for r in recs} == {"-"}
