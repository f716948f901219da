"""Edge cases of the per-read path on the GPU against the oracle pipeline: empty batch, empty / tiny / all-N reads,
N runs inside a read, a read made of two contigs, a long read (70 kb: multi-band fill, wide edit-distance classes),
high-divergence reads that the divergence filter has to look at exactly, duplicated reads, lower-case input."""
import numpy as np
import pytest

import oracle
import oracle.pipeline as pl
import synth

pytestmark = pytest.mark.gpu


def _check(vb, gpu_ctx, ref, reads, mode="H", **kw):
    opt = vb.default_option(mode, **kw)
    ix = vb.Index(ref, ctx=gpu_ctx)
    ox = oracle.Index(ref)
    ctg = pl.Contigs([n for n, _ in ref], [s for _, s in ref])
    outs = [vb.Aligner(ix, opt, mode, workers=1).align_batch(reads), vb.Aligner(ix, opt, mode, workers=3, chunk_reads=2).align_batch(reads)]
    n_rec = 0
    for got in outs:
        for (rid, seq), g in zip(reads, got):
            want = pl.align_read(rid, seq.upper(), ox, ctg, opt, mode)
            assert [tuple(r) for r in g] == [tuple(w) for w in want], rid
            n_rec += len(g)
    ix.close()
    return n_rec


def test_empty_batch_and_degenerate_reads(gpu_ctx):
    import vacmap_b200 as vb
    ref = synth.make_reference(61, 150000, n_contigs=2)
    ix = vb.Index(ref, ctx=gpu_ctx)
    al = vb.Aligner(ix, vb.default_option("H"), "H")
    assert al.align_batch([]) == []
    off, recs, cig = al.align_packed(b"", np.zeros(1, np.int64))
    assert list(off) == [0] and len(recs) == 0 and len(cig) == 0
    ix.close()
    rng = np.random.default_rng(3)
    good = synth.make_reads(ref, 62, 3, read_len=4000, err=0.08)
    reads = [("empty", ""), ("tiny", "ACGTACG"), ("k-1", "ACGTACGTACGTAC"), ("allN", "N" * 3000),
             ("random", synth.random_seq(rng, 5000).tobytes().decode())] + good + [("dup", good[0][1]), ("lower", good[1][1].lower())]
    assert _check(vb, gpu_ctx, ref, reads) >= 5


def test_reads_with_N_runs_and_contig_junction(gpu_ctx):
    import vacmap_b200 as vb
    ref = synth.make_reference(63, 200000, n_contigs=2)
    reads = synth.make_reads(ref, 64, 6, read_len=7000, err=0.08)
    out = []
    for i, (rid, s) in enumerate(reads):
        s = list(s)
        for p in range(500 + 137 * i, len(s), 1500):        # N runs of 1..40 bases
            for q in range(p, min(len(s), p + 1 + (i * 13) % 40)):
                s[q] = "N"
        out.append((rid + "_N", "".join(s)))
    (n1, s1), (n2, s2) = ref[0], ref[1]
    out.append(("junction", s1[-3500:] + s2[:3500]))          # chimeric across the contig boundary
    comp = str.maketrans("ACGTN", "TGCAN")
    out.append(("junction_rc", (s1[-3000:] + s2[:4000]).translate(comp)[::-1]))
    for eqx in (False, True):
        assert _check(vb, gpu_ctx, ref, out, eqx=eqx) >= 6


def test_long_read_and_divergent_reads(gpu_ctx):
    import vacmap_b200 as vb
    ref = synth.make_reference(65, 400000)
    long_read = synth.make_reads(ref, 66, 1, read_len=70000, err=0.10)
    # a big deletion in the read: the fill sees a ~300 x several-kb segment (multi-band class)
    name, s = ref[0]
    deleted = s[100000:104000] + s[110000:114000]
    noisy = synth.make_reads(ref, 67, 4, read_len=6000, err=0.18) + synth.make_reads(ref, 68, 3, read_len=6000, err=0.24)
    reads = long_read + [("del6k", deleted)] + noisy
    assert _check(vb, gpu_ctx, ref, reads) >= 3
    assert _check(vb, gpu_ctx, ref, noisy, mode="S") >= 1


def test_a_read_that_cannot_be_processed_is_dropped_alone(gpu_ctx, monkeypatch):
    """The reference's worker loses only the read that raised (`except Exception: continue`, clrnano:24116-24125).
    VM_TEST_FAIL_LEN injects a failure for reads of one length: that read comes back without records, every other
    read of the batch -- same chunk or not -- is unaffected."""
    import vacmap_b200 as vb
    ref = synth.make_reference(71, 200000)
    reads = synth.make_reads(ref, 72, 9, read_len=5000, err=0.08)
    bad = ("bad", reads[0][1][:4321])
    batch = reads[:4] + [bad] + reads[4:]
    opt = vb.default_option("H")
    ix = vb.Index(ref, ctx=gpu_ctx)
    want = vb.Aligner(ix, opt, "H", workers=1).align_batch(batch)
    assert want[4]                                   # it maps when nothing is injected
    monkeypatch.setenv("VM_TEST_FAIL_LEN", "4321")
    for workers, chunk in ((1, 0), (3, 4)):
        got = vb.Aligner(ix, opt, "H", workers=workers, chunk_reads=chunk).align_batch(batch)
        assert got[4] == []
        assert got[:4] == want[:4] and got[5:] == want[5:]
    ix.close()
