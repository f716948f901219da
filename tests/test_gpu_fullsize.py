"""BASELINE configs[1] at full size (10 000 reads x 15 kb, 10 % error, 5 Mb reference, mode H) through size-independent
properties, plus oracle parity on a random sample of the very same batch:
  * every CIGAR consumes exactly its read (clips included) and exactly its reference span;
  * the pipelined path (6 workers, jobs in flight) returns bit-identical arrays to the lock-step path;
  * 48 reads drawn from the batch equal the oracle pipeline record for record (CIGAR, MAPQ, coordinates)."""
import numpy as np
import pytest

import oracle
import oracle.pipeline as pl
import synth

pytestmark = pytest.mark.gpu

Q_OPS = np.array([1, 1, 0, 0, 1, 0, 0, 1, 1], dtype=np.int64)   # MIDNSHP=X: consumes query
R_OPS = np.array([1, 0, 1, 1, 0, 0, 0, 1, 1], dtype=np.int64)   # consumes reference


@pytest.fixture(scope="module")
def workload():
    ref = synth.make_reference(1, 5_000_000)
    reads = synth.make_reads(ref, 11, 10000, read_len=15000, err=0.10)
    enc = [s.encode() for _, s in reads]
    off = np.concatenate([[0], np.cumsum([len(e) for e in enc])]).astype(np.int64)
    return ref, reads, b"".join(enc), off


def test_full_size_properties_and_sampled_parity(gpu_ctx, workload):
    import vacmap_b200 as vb
    ref, reads, cat, off = workload
    opt = vb.default_option("H")
    ix = vb.Index(ref, w=10, k=15, ctx=gpu_ctx)
    piped = vb.Aligner(ix, opt, "H")                 # default: pipelined workers
    h1 = piped.submit_packed(cat, off)
    h2 = piped.submit_packed(cat, off)               # two jobs in flight
    a = piped.wait(h1)
    a2 = piped.wait(h2)
    b = vb.Aligner(ix, opt, "H", workers=1).align_packed(cat, off)
    for x, y, z in zip(a, b, a2):
        assert x.shape == y.shape and (x == y).all() and (x == z).all()
    rec_off, recs, cig = a
    n_mapped = int((np.diff(rec_off) > 0).sum())
    assert n_mapped >= 0.99 * len(reads)             # simulated reads all come from the reference
    # CIGAR bookkeeping of every record
    op, ln = (cig & 0xf).astype(np.int64), (cig >> 4).astype(np.int64)
    qc = np.concatenate([[0], np.cumsum(ln * Q_OPS[op])])
    rc = np.concatenate([[0], np.cumsum(ln * R_OPS[op])])
    lo, hi = recs["cigar_off"], recs["cigar_off"] + recs["cigar_len"]
    read_of = np.repeat(np.arange(len(reads)), np.diff(rec_off))
    read_len = np.diff(off)[read_of]
    assert ((qc[hi] - qc[lo]) == read_len).all()                      # soft clips included: the whole read
    assert ((rc[hi] - rc[lo]) == (recs["r_en"] - recs["r_st"])).all()
    assert (recs["q_st"] >= 0).all() and (recs["q_en"] <= read_len).all() and (recs["q_st"] < recs["q_en"]).all()
    assert ((recs["mapq"] >= 0) & (recs["mapq"] <= 60)).all()
    first_is_clip = (op[lo] == 4)
    assert (np.where(first_is_clip, ln[lo], 0) == recs["q_st"]).all()
    # oracle parity on a sample of the same batch
    ox = oracle.Index(ref, w=10, k=15)
    ctg = pl.Contigs([n for n, _ in ref], [s for _, s in ref])
    rng = np.random.default_rng(7)
    for i in sorted(rng.choice(len(reads), 48, replace=False)):
        rid, seq = reads[i]
        want = [tuple(w) for w in pl.align_read(rid, seq, ox, ctg, opt, "H")]
        got = [tuple(r) for r in piped.rows_of(rid, recs[rec_off[i]:rec_off[i + 1]], cig)]
        assert got == want, rid
    ix.close()
