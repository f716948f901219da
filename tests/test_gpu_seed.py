"""GPU parity of the seeding stage (chunk-parallel sketch, hash lookup, cluster filter, strand flip)
against the oracle's restatement of `vacmap_index.map` + get_reversed_chain_numpy_rough."""
import numpy as np
import pytest

import oracle
import oracle.pipeline as pl
import synth

pytestmark = pytest.mark.gpu


def with_ns(rng, s, n_runs):
    a = np.frombuffer(s.encode(), dtype=np.uint8).copy()
    for _ in range(n_runs):
        p = int(rng.integers(0, len(a) - 1))
        ln = int(rng.choice([1, 1, 2, 5, 20, 60, 300]))
        a[p:p + ln] = ord("N")
    return a.tobytes().decode()


def check(ix, ox, reads, check_num):
    from vacmap_b200.align import seed_batch
    got = seed_batch(ix, reads, check_num)
    for s, (rows, rev) in zip(reads, got):
        want = ox.map(s.upper(), check_num=check_num, mid_occ=-1)
        wrev, want = pl.reverse_rough(want, len(s))
        assert rev == wrev
        assert rows.shape == want.shape and (rows == want).all(), len(s)


def test_seeding_matches_oracle(gpu_ctx):
    import vacmap_b200 as vb
    rng = np.random.default_rng(41)
    ref = synth.make_reference(41, 600000, n_contigs=2, repeat_frac=0.15)
    ix, ox = vb.Index(ref, ctx=gpu_ctx), oracle.Index(ref)
    reads = [s for _, s in synth.make_reads(ref, 42, 40, read_len=7000, err=0.08)]
    reads += [with_ns(rng, s, int(rng.integers(1, 12))) for s in reads[:16]]          # ambiguous bases, incl. in warm-up zones
    reads += ["ACGT", "A" * 500, "ACGTACGTACGTACGTACGTACGTACGTACGTACGT" * 20, "N" * 300, reads[0][:14], reads[0][:15], reads[0][:25],
              reads[1][:127], reads[1][:128], reads[1][:129], reads[2][:256 + 24]]
    palin = "ACGTTGCATGCAACGT" * 40                                                     # rich in symmetric k-mers (k = 15?) and repeats
    reads += [palin, reads[3][:1000] + palin + reads[3][1000:2000]]
    check(ix, ox, reads, 100)
    check(ix, ox, reads, -1)
    check(ix, ox, reads[:30], 3)        # the top-N cluster filter actually drops clusters
    ix.close()


def test_seeding_other_kw(gpu_ctx):
    import vacmap_b200 as vb
    ref = synth.make_reference(43, 300000)
    reads = [s for _, s in synth.make_reads(ref, 44, 12, read_len=5000, err=0.01, ratio=(1, 1, 1))]
    for w, k in ((10, 19), (5, 11), (19, 21)):
        ix, ox = vb.Index(ref, w=w, k=k, ctx=gpu_ctx), oracle.Index(ref, w=w, k=k)
        check(ix, ox, reads, 100)
        ix.close()


def test_local_reseeding_matches_oracle(gpu_ctx):
    """Stage-level local re-seeding (vm_local_reseed_batch): for the guide chains the oracle pipeline selects on real
    reads (SVs, both strands, two contigs), the anchors of every guide job -- windows and guide points built by the
    oracle's restatement of :23095-23136 -- must equal orc_local_reseed's, in the reference's emission order."""
    import vacmap_b200 as vb
    from vacmap_b200.align import local_reseed_batch
    ref = synth.make_reference(45, 400000, n_contigs=2, repeat_frac=0.10)
    reads = synth.make_reads(ref, 46, 14, read_len=8000, err=0.10, sv_frac=0.5)
    ox = oracle.Index(ref)
    ctg = pl.Contigs([n for n, _ in ref], [s for _, s in ref])
    opt = vb.default_option("S")           # mode S re-seeds every secondary chain
    oriented, jobs, want = [], [], []
    for rid, seq in reads:
        seq = seq.upper()
        L = len(seq)
        mapq, scores, path_list = pl.decode_hit(ox, seq, L, ox.k, opt, "S")
        if scores == 0.0:
            continue
        rc_seq = pl.revcomp(seq)
        if scores < 0.0:
            seq, rc_seq = rc_seq, seq
        chains = [np.array(p, dtype=np.int64) for p in path_list]
        chains = pl.drop_somechains(pl.merge_chain(chains))
        ri = len(oriented)
        oriented.append(seq)
        for ch in chains:
            wins, raw = pl.guide_windows(ch, ctg)
            readstart = max(0, int(raw[0][0]) - 7000)
            readend = min(L - 9 + 1, int(raw[-1][0]) + 7000)
            jobs.append((ri, wins, raw, readstart, readend))
            want.append(oracle.local_reseed_scan(ctg, wins, raw, seq, rc_seq, 9, readstart, readend))
    assert len(jobs) >= 10 and sum(len(w) for w in want) > 5000
    ix = vb.Index(ref, ctx=gpu_ctx)
    got = local_reseed_batch(ix, oriented, jobs)
    for (ri, wins, raw, a, b), g, w in zip(jobs, got, want):
        assert g.shape == w.shape and (g == w).all(), (ri, len(wins), len(raw))
    ix.close()


def test_local_reseeding_matches_reference_guide_1(gpu_ctx):
    """Stage golden from the REFERENCE's own get_localmap_multi_all_forDP_inv_guide_1 (clrnano:23069-23345, run in the
    build container, tests/golden/guide1.npz): vm_local_reseed_batch must hand back the same anchors in the same order
    for the same guide chains (windows built by the oracle's restatement of :23095-23136, the host glue's job)."""
    import guide1_cases
    import vacmap_b200 as vb
    from vacmap_b200.align import local_reseed_batch
    by_case = {}
    for j in guide1_cases.jobs():
        by_case.setdefault(j["case"], []).append(j)
    n = 0
    for name, js in by_case.items():
        ctg = pl.Contigs([c for c, _ in js[0]["ref"]], [s for _, s in js[0]["ref"]])
        k, w = (15, 10)
        ix = vb.Index(js[0]["ref"], w=w, k=k, ctx=gpu_ctx)
        oriented, jobs = [], []
        for j in js:
            wins, raw = pl.guide_windows(j["chain"], ctg)
            oriented.append(j["seq"])
            jobs.append((len(oriented) - 1, wins, raw, j["range"][0], j["range"][1]))
        got = local_reseed_batch(ix, oriented, jobs)
        for j, g in zip(js, got):
            assert g.shape == j["out"].shape and (g == j["out"]).all()
            n += len(g)
        ix.close()
    assert n > 10000


def test_device_built_index_equals_host_built_index(gpu_ctx, monkeypatch, tmp_path):
    """the index built on the device (chunk-parallel sketch of the contigs, stable radix sort, CAS hash table,
    radix-sorted 9-mers) must be the index the single-threaded host build of round 1 makes: same keys, counts and
    occurrence order, same default occurrence cap, same seeds and same records; and it survives a trip through a
    minimap2 `.mmi` file (vacmap:324-344)."""
    import vacmap_b200 as vb
    from vacmap_b200 import mmi
    from vacmap_b200.align import seed_batch
    ref = synth.make_reference(5, 300000, n_contigs=3, repeat_frac=0.1)
    ref.append(("withN", "ACGT" * 50 + "N" * 30 + "ACGTTGCA" * 40 + "nnnnacgtacgatcgatcgatcgatcgtagctagctagctagcatcgatcgatcga" * 5))
    reads = synth.make_reads(ref[:3], 77, 12, read_len=5000, err=0.08, sv_frac=0.5)
    ix = vb.Index(ref, w=10, k=15, ctx=gpu_ctx)
    dev = ix.minimizers()
    seeds = seed_batch(ix, [s for _, s in reads])
    recs = vb.Aligner(ix, vb.default_option("H"), "H").align_batch(reads)
    monkeypatch.setenv("VM_INDEX_HOST", "1")
    ixh = vb.Index(ref, w=10, k=15, ctx=gpu_ctx)
    monkeypatch.delenv("VM_INDEX_HOST")
    host = ixh.minimizers()
    assert (ix.n_keys, ix.n_minimizers, ix.mid_occ) == (ixh.n_keys, ixh.n_minimizers, ixh.mid_occ)
    for a, b in zip(dev, host):
        assert a.shape == b.shape and (a == b).all()
    assert all(ix.seq(n) == ixh.seq(n) for n, _ in ref)
    for (a, fa), (b, fb) in zip(seeds, seed_batch(ixh, [s for _, s in reads])):
        assert fa == fb and a.shape == b.shape and (a == b).all()
    assert recs == vb.Aligner(ixh, vb.default_option("H"), "H").align_batch(reads)
    ox = oracle.Index(ref)
    assert (ix.n_keys, ix.n_minimizers, ix.mid_occ) == (ox.n_keys, ox.n_occ, ox.mid_occ_default)
    # through a .mmi file: what the writer stores is what the reader finds, and an index opened from it is the same index
    p = str(tmp_path / "ref.fa.w10_k15.mmi")
    ix.write_mmi(p)
    m = mmi.read_mmi(p, with_minimizers=True)
    assert m["names"] == [n for n, _ in ref] and m["seqs"] == [ix.seq(n) for n, _ in ref]
    assert (m["keys"] == dev[0]).all() and (m["counts"] == dev[1]).all() and (m["occ"] == dev[2]).all()
    ix2 = vb.Index(p, w=10, k=15, ctx=gpu_ctx)
    assert (ix2.n_keys, ix2.n_minimizers, ix2.mid_occ) == (ix.n_keys, ix.n_minimizers, ix.mid_occ)
    assert recs == vb.Aligner(ix2, vb.default_option("H"), "H").align_batch(reads)
    # a second handle over the same device arrays (what a rank does with an index another rank broadcast)
    arrs, meta = ix.arrays()
    ix3 = vb.Index.adopt(ix.names, ix.lens, [p_ for p_, _ in arrs], [b_ for _, b_ in arrs], meta, ctx=gpu_ctx)
    assert recs == vb.Aligner(ix3, vb.default_option("H"), "H").align_batch(reads)
    for x in (ix3, ix2, ixh, ix):
        x.close()
