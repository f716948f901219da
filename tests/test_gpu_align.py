"""GPU parity of the base-level kernels (edit distance, z-drop extension, global fill) against the
oracle's restated natives, on raw sequence pairs through the stage-level C-ABI entry point."""
import numpy as np
import pytest

import oracle
import synth

pytestmark = pytest.mark.gpu


def pairs(rng, n, lo, hi, err=0.12, unrelated_tail=False):
    ts, qs = [], []
    for _ in range(n):
        t = synth.random_seq(rng, int(rng.integers(lo, hi)))
        q = synth.mutate(rng, t, err)
        if rng.random() < 0.2:   # a long indel inside
            p = int(rng.integers(0, max(1, len(q) - 1)))
            q = np.concatenate([q[:p], synth.random_seq(rng, int(rng.integers(20, 200))), q[p:]]) if rng.random() < 0.5 \
                else np.concatenate([q[:p], q[p + int(rng.integers(20, 120)):]])
        if unrelated_tail:
            cut = int(rng.integers(0, len(q) + 1))
            q = np.concatenate([q[:cut], synth.random_seq(rng, int(rng.integers(50, 400)))])
        if len(q) == 0:
            q = synth.random_seq(rng, 3)
        ts.append(t.tobytes().decode())
        qs.append(q.tobytes().decode())
    return ts, qs


def test_fill_cigars_match_oracle(gpu_ctx):
    from vacmap_b200.align import pairs_batch
    rng = np.random.default_rng(31)
    ts, qs = pairs(rng, 200, 5, 420)
    t2, q2 = pairs(rng, 12, 500, 1400)          # targets beyond one 256-row band
    ts += t2 + ["A", "ACGT", "ACGTNNACGT", "TTTTTTTTTT"]
    qs += q2 + ["ACGTACGT", "A", "ACGTNNACGA", "TTTTT"]
    # the common case in bulk, so that every slot class (two pairs per warp up to 64 band rows, one beyond) sees
    # full warps, a half-empty last warp and partners of unequal length
    for i in range(523):
        n = int(rng.integers(96, 760))
        t = synth.random_seq(rng, n)
        q = synth.mutate(rng, t, float(rng.choice([0.05, 0.10, 0.12])))[:760]
        ts.append(t.tobytes().decode())
        qs.append(q.tobytes().decode())
    for eqx in (False, True):
        got = pairs_batch("fill", ts, qs, eqx=eqx, ctx=gpu_ctx)
        for t, q, g in zip(ts, qs, got):
            assert g == oracle.k_cigar(t, q, 2, -4, 4, 2, 24, 1, -1, -1, eqx)[0], (len(t), len(q))


def test_extension_endpoints_match_oracle(gpu_ctx):
    from vacmap_b200.align import pairs_batch
    rng = np.random.default_rng(32)
    ts, qs = pairs(rng, 150, 5, 900, unrelated_tail=True)
    ts += ["ACGTACGTAC", "A" * 50]
    qs += ["TTTTTTTTTT", "A" * 70]
    got = pairs_batch("extend", ts, qs, ctx=gpu_ctx)
    for t, q, g in zip(ts, qs, got):
        r = oracle.k_cigar(t, q, 2, -4, 4, 4, 4, 4, 100, 50)
        assert g == (r[2], r[3]), (len(t), len(q))


def test_edit_distance_matches_oracle(gpu_ctx):
    from vacmap_b200.align import pairs_batch
    rng = np.random.default_rng(33)
    ts, qs = pairs(rng, 120, 1, 700, err=0.2)
    t2, q2 = pairs(rng, 6, 3000, 9000, err=0.15)
    ts += t2 + ["ACGT"]
    qs += q2 + ["ACGT"]
    got = pairs_batch("distance", ts, qs, ctx=gpu_ctx)
    for t, q, g in zip(ts, qs, got):
        assert g == oracle.edit_distance(q, t), (len(t), len(q))


def test_banded_edit_distance_is_exact_inside_the_band(gpu_ctx):
    """kind 3: exact when the distance is <= k, some value > k otherwise (slot reuse: bands far narrower than
    the pattern, bands of a few blocks, k = 0, k one below / equal to the true distance)."""
    from vacmap_b200.align import pairs_batch
    rng = np.random.default_rng(34)
    ts, qs = pairs(rng, 60, 1, 900, err=0.2)
    t2, q2 = pairs(rng, 10, 6000, 16000, err=0.12)
    t3, q3 = pairs(rng, 4, 20000, 40000, err=0.03)
    ts += t2 + t3 + ["ACGT", "ACGTACGTAA"]
    qs += q2 + q3 + ["ACGT", "ACGTACGTCC"]
    exact = pairs_batch("distance", ts, qs, ctx=gpu_ctx)
    for t, q, d in zip(ts[:70], qs[:70], exact[:70]):
        assert d == oracle.edit_distance(q, t)
    for mk in (lambda d, m: d, lambda d, m: max(d - 1, 0), lambda d, m: int(0.2 * m), lambda d, m: d + 70,
               lambda d, m: 0, lambda d, m: int(0.5 * m), lambda d, m: d // 2):
        band = [mk(d, min(len(t), len(q))) for t, q, d in zip(ts, qs, exact)]
        got = pairs_batch("distance", ts, qs, ctx=gpu_ctx, band=band)
        for t, q, d, k, g in zip(ts, qs, exact, band, got):
            if d <= k:
                assert g == d, (len(t), len(q), d, k, g)
            else:
                assert g > k, (len(t), len(q), d, k, g)


def test_banded_fill_certificate_and_fallback(gpu_ctx):
    """Pairs inside the banded kernel's size range (96..768) whose optimum hugs or leaves the band: unrelated
    sequences, heavy divergence, a large insertion / deletion, tandem shifts.  Whether the certificate holds or
    the pair falls back to the full-matrix kernel, the CIGAR must be the oracle's (full-matrix) CIGAR."""
    from vacmap_b200.align import pairs_batch
    rng = np.random.default_rng(35)
    ts, qs = [], []
    for i in range(160):
        n = int(rng.integers(100, 700))
        t = synth.random_seq(rng, n)
        kind = i % 8
        if kind == 0:
            q = synth.random_seq(rng, int(rng.integers(100, 700)))                     # unrelated
        elif kind == 1:
            q = synth.mutate(rng, t, 0.30)                                             # very divergent
        elif kind == 2:
            p = int(rng.integers(10, n - 10))
            q = np.concatenate([t[:p], synth.random_seq(rng, int(rng.integers(40, 300))), t[p:]])[:760]   # big insertion
        elif kind == 3:
            p = int(rng.integers(10, n // 2))
            q = np.concatenate([t[:p], t[p + int(rng.integers(40, n // 2)):]])          # big deletion
        elif kind == 4:
            k = int(rng.integers(5, 60))
            q = np.concatenate([t[k:], t[:k]])                                         # rotation: shifted diagonal
        elif kind == 5:
            q = synth.mutate(rng, t, 0.10)
            q[rng.integers(0, len(q), size=len(q) // 15)] = ord("N")                    # N runs
        else:
            q = synth.mutate(rng, t, float(rng.choice([0.02, 0.10, 0.15, 0.20])))
        if len(q) < 96:
            q = np.concatenate([q, synth.random_seq(rng, 96)])
        ts.append(t.tobytes().decode())
        qs.append(q.tobytes().decode())
    # the common case in bulk, so that every slot class (two pairs per warp up to 64 band rows, one beyond) sees
    # full warps, a half-empty last warp and partners of unequal length
    for i in range(523):
        n = int(rng.integers(96, 760))
        t = synth.random_seq(rng, n)
        q = synth.mutate(rng, t, float(rng.choice([0.05, 0.10, 0.12])))[:760]
        ts.append(t.tobytes().decode())
        qs.append(q.tobytes().decode())
    for eqx in (False, True):
        got = pairs_batch("fill", ts, qs, eqx=eqx, ctx=gpu_ctx)
        for t, q, g in zip(ts, qs, got):
            assert g == oracle.k_cigar(t, q, 2, -4, 4, 2, 24, 1, -1, -1, eqx)[0], (len(t), len(q))
