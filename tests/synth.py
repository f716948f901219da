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


# ---------------------------------------------------------------------------
# sequences: references, reads with ONT/HiFi-like errors, simple SV donors
# ---------------------------------------------------------------------------
_B = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


def random_seq(rng, n):
    return _B[rng.integers(0, 4, size=n)]


def make_reference(seed, length, n_contigs=1, repeat_frac=0.05):
    """i.i.d. ACGT contigs `chr1..`, with `repeat_frac` of the sequence overwritten by diverged copies
    of earlier 0.3-6 kb segments (exercises the occurrence filter / coverage logic)."""
    rng = np.random.default_rng(seed)
    per = length // n_contigs
    contigs = []
    for c in range(n_contigs):
        s = random_seq(rng, per).copy()
        covered = 0
        while covered < repeat_frac * per and per > 20000:
            ln = int(rng.integers(300, 6000))
            src = int(rng.integers(0, per - ln))
            dst = int(rng.integers(0, per - ln))
            seg = s[src:src + ln].copy()
            nmut = int(ln * rng.uniform(0, 0.1))
            pos = rng.integers(0, ln, size=nmut)
            seg[pos] = _B[rng.integers(0, 4, size=nmut)]
            s[dst:dst + ln] = seg
            covered += ln
        contigs.append(("chr%d" % (c + 1), s.tobytes().decode()))
    return contigs


def mutate(rng, seq_u8, err, ratio=(4, 3, 3)):
    """per-base i.i.d. errors at rate `err`, sub:ins:del = ratio"""
    n = len(seq_u8)
    r = rng.random(n)
    tot = float(sum(ratio))
    ps, pi = err * ratio[0] / tot, err * ratio[1] / tot
    kinds = np.where(r < ps, 1, np.where(r < ps + pi, 2, np.where(r < err, 3, 0)))
    rnd = rng.integers(0, 4, size=n)
    # vectorised assembly: kind 0 copy, 1 substitution, 2 copy + inserted base, 3 deletion
    sub = _B[rnd]
    sub = np.where(sub != seq_u8, sub, _B[(rnd + 1) & 3])
    first = np.where(kinds == 1, sub, seq_u8)
    counts = np.where(kinds == 3, 0, np.where(kinds == 2, 2, 1))
    pos = np.cumsum(counts) - counts
    out = np.empty(int(counts.sum()), dtype=np.uint8)
    keep = kinds != 3
    out[pos[keep]] = first[keep]
    ins = kinds == 2
    out[pos[ins] + 1] = _B[rnd[ins]]
    return out


def make_reads(contigs, seed, n_reads, read_len=15000, err=0.10, ratio=(4, 3, 3), sv_frac=0.0):
    """Reads drawn uniformly; strand ~ Bernoulli(0.5).  With `sv_frac`, a read gets one event from
    {DEL, INS, INV, DUP, TRA} of 100-1000 bp applied to its source segment before errors."""
    rng = np.random.default_rng(seed)
    arrs = [np.frombuffer(s.encode(), dtype=np.uint8) for _, s in contigs]
    reads = []
    for i in range(n_reads):
        c = int(rng.integers(0, len(arrs)))
        ref = arrs[c]
        ln = min(read_len, len(ref) - 1)
        st = int(rng.integers(0, len(ref) - ln))
        seg = ref[st:st + ln].copy()
        if rng.random() < sv_frac and ln > 4000:
            sz = int(rng.integers(100, 1000))
            p = int(rng.integers(1000, ln - 1000 - sz))
            kind = int(rng.integers(0, 5))
            if kind == 0:
                seg = np.concatenate([seg[:p], seg[p + sz:]])
            elif kind == 1:
                seg = np.concatenate([seg[:p], random_seq(rng, sz), seg[p:]])
            elif kind == 2:
                seg = np.concatenate([seg[:p], _COMP[seg[p:p + sz]][::-1], seg[p + sz:]])
            elif kind == 3:
                seg = np.concatenate([seg[:p + sz], seg[p:p + sz], seg[p + sz:]])
            else:
                q = int(rng.integers(0, len(ref) - sz))
                seg = np.concatenate([seg[:p], ref[q:q + sz], seg[p:]])
        if rng.random() < 0.5:
            seg = _COMP[seg][::-1]
        seg = mutate(rng, seg, err, ratio)
        reads.append(("read_%d" % i, seg.tobytes().decode()))
    return reads


def asm_e2e_inputs():
    """The asm end-to-end fixture's inputs (tests/golden/make_golden.py::asm_e2e_inputs, same seeds): a 1.3 Mb 2-contig
    reference and one 520 kb contig read cut from it with a 4 kb inversion, a 2.5 kb deletion, a 1.2 kb insertion and
    0.5 % divergence."""
    ref = make_reference(91, 1300000, n_contigs=2)
    rng = np.random.default_rng(92)
    src = np.frombuffer(ref[0][1].encode(), dtype=np.uint8)[40000:563700].copy()
    comp = np.zeros(256, np.uint8)
    for x, y in zip(b"ACGT", b"TGCA"):
        comp[x] = y
    parts = [src[:150000], comp[src[150000:154000]][::-1], src[154000:300000], src[302500:420000],
             random_seq(rng, 1200), src[420000:]]
    read = mutate(rng, np.concatenate(parts), 0.005, ratio=(1, 1, 1))
    return ref, read.tobytes().decode()


def asm_e2e_inputs_2():
    """Second asm end-to-end fixture (tests/golden/make_golden.py::gen_asm_e2e_2, same seeds): a 1.6 Mb 2-contig reference
    and a 560 kb contig read taken from chr1's REVERSE strand, carrying a 30 kb piece of chr2 (a translocation), a 6 kb
    tandem duplication and a 3 kb deletion, 0.3 % divergence."""
    ref = make_reference(93, 1600000, n_contigs=2)
    rng = np.random.default_rng(94)
    c1 = np.frombuffer(ref[0][1].encode(), dtype=np.uint8)
    c2 = np.frombuffer(ref[1][1].encode(), dtype=np.uint8)
    src = c1[100000:640000]
    parts = [src[:120000], c2[300000:330000], src[120000:260000], src[254000:260000], src[260000:400000], src[403000:]]
    fwd = np.concatenate(parts)
    read = mutate(rng, _COMP[fwd][::-1].copy(), 0.003, ratio=(1, 1, 1))
    return ref, read.tobytes().decode()
