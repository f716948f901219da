"""Deterministic synthetic inputs shared by the tests, bench and golden generator.

Anchor sets mimic what the seeding stage produces for a long noisy read:
a colinear true chain (optionally with an inverted / translocated segment),
repeat copies sharing read positions, and uniformly random noise anchors --
tie-heavy on purpose (many equal scores), because tie-breaking is the hard
part of parity (hard part 1).
"""
import numpy as np


def anchors_global(rng, L=15000, k=15, n_true=550, n_noise=1500, ref_len=5_000_000,
                   sv=True, repeats=True, dup_frac=0.05):
    """int64[n,4] rows (readpos, refpos, strand, len); NOT sorted."""
    rows = []
    xs = np.sort(rng.choice(L - k, size=min(n_true, L - k), replace=False))
    r0 = int(rng.integers(20000, ref_len - 3 * L - 20000))
    strand = 1 if rng.random() < 0.5 else -1
    inv_lo, inv_hi = (L // 3, L // 3 + int(rng.integers(500, 4000))) if sv and rng.random() < 0.5 else (-1, -1)
    drift = 0
    for x in xs:
        if rng.random() < 0.02:
            drift += int(rng.integers(-20, 21))
        s = strand
        if inv_lo <= x < inv_hi:
            s = -strand
        if s == 1:
            y = r0 + x + drift
        else:
            y = r0 + 2 * L - x - k + drift
        rows.append((int(x), int(y), s, k))
        if repeats and rng.random() < dup_frac:
            for _ in range(int(rng.integers(1, 6))):
                rows.append((int(x), int(rng.integers(0, ref_len)), 1 if rng.random() < 0.5 else -1, k))
    for _ in range(n_noise):
        rows.append((int(rng.integers(0, L - k)), int(rng.integers(0, ref_len)),
                     1 if rng.random() < 0.5 else -1, k))
    a = np.array(rows, dtype=np.int64)
    return a[rng.permutation(len(a))]


def anchors_tieheavy(rng, n=400, L=2000, k=15, ref_len=20000):
    """Small, extremely tie-heavy set: few distinct positions, tandem-like."""
    x = rng.integers(0, L // 16, size=n) * 16
    y = rng.integers(0, ref_len // 16, size=n) * 16
    s = rng.choice([-1, 1], size=n)
    ln = np.full(n, k)
    return np.stack([x, y, s, ln], axis=1).astype(np.int64)


def anchors_local(rng, L=15000, k=9, n_true=1000, n_noise=300, ref0=1_000_000, multi=False):
    """Local-stage style anchors: len in [9, 19+], sorted later by read END."""
    rows = []
    x = int(rng.integers(0, 50))
    drift = 0
    strand = 1
    while x < L - 40 and len(rows) < n_true:
        ln = int(rng.integers(k, 20))
        if rng.random() < 0.1:
            drift += int(rng.integers(-8, 9))
        if multi and rng.random() < 0.003:
            strand = -strand
            drift += int(rng.integers(-3000, 3000))
        if strand == 1:
            y = ref0 + x + drift
        else:
            y = ref0 + 2 * L - x - ln + drift
        rows.append((x, y, strand, ln))
        x += ln + int(rng.integers(-4, 12))
        x = max(x, 0)
    for _ in range(n_noise):
        ln = int(rng.integers(k, 14))
        rows.append((int(rng.integers(0, L - 20)), ref0 + int(rng.integers(-7000, L + 7000)),
                     1 if rng.random() < 0.5 else -1, ln))
    a = np.array(rows, dtype=np.int64)
    return a[rng.permutation(len(a))]
