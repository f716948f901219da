"""Self-consistency of the oracle's restated natives (seeding, k_cigar, edit distance) -- the stages whose
third-party sources are absent from the reference tree (parity unpinned)."""
import numpy as np

import oracle
import synth


def brute_minimizers(seq, w, k):
    """Definition check on clean ACGT input: position p is emitted iff its canonical-k-mer hash equals the
    minimum over some full window of w consecutive k-mers (symmetric k-mers excluded)."""
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    mask = (1 << 2 * k) - 1

    def h64(key):
        key = (~key + (key << 21)) & mask
        key ^= key >> 24
        key = (key + (key << 3) + (key << 8)) & mask
        key ^= key >> 14
        key = (key + (key << 2) + (key << 4)) & mask
        key ^= key >> 28
        key = (key + (key << 31)) & mask
        return key
    enc = {"A": 0, "C": 1, "G": 2, "T": 3}
    hs = []
    for i in range(len(seq) - k + 1):
        km = seq[i:i + k]
        rc = "".join(comp[c] for c in reversed(km))
        f = int("".join(str(enc[c]) for c in km), 4)
        r = int("".join(str(enc[c]) for c in rc), 4)
        assert f != r
        hs.append(h64(min(f, r)))
    out = set()
    for s in range(len(hs) - w + 1):
        m = min(hs[s:s + w])
        for j in range(s, s + w):
            if hs[j] == m:
                out.add(j + k - 1)
    return sorted(out)


def test_sketch_matches_window_minimum_definition():
    rng = np.random.default_rng(4)
    for _ in range(5):
        seq = synth.random_seq(rng, 600).tobytes().decode()
        h, y = oracle.sketch(seq, 10, 15)
        assert sorted(int(v) >> 1 for v in y) == brute_minimizers(seq, 10, 15)


def test_map_finds_the_true_diagonal():
    ref = synth.make_reference(9, 100000, repeat_frac=0.0)
    ix = oracle.Index(ref)
    read = ref[0][1][20000:26000]
    a = ix.map(read, check_num=100)
    good = (a[:, 2] == 1) & (a[:, 1] - a[:, 0] == 20000)
    assert good.sum() > 0.9 * len(a) and len(a) > 500
    rc = oracle.pipeline_revcomp(read) if hasattr(oracle, "pipeline_revcomp") else read.translate(str.maketrans("ACGT", "TGCA"))[::-1]
    b = ix.map(rc, check_num=100)
    assert (b[:, 2] == -1).sum() > 0.9 * len(b)


def test_edit_distance_bitvector_equals_dp():
    rng = np.random.default_rng(5)
    for _ in range(40):
        n, m = int(rng.integers(1, 400)), int(rng.integers(1, 400))
        a = synth.random_seq(rng, n)
        b = synth.mutate(rng, a, 0.2)[:m] if rng.random() < 0.7 else synth.random_seq(rng, m)
        a, b = a.tobytes().decode(), b.tobytes().decode()
        if not b:
            continue
        assert oracle.edit_distance(a, b) == oracle.edit_distance_dp(a, b)
    assert oracle.edit_distance("kitten", "sitting") == 3


def cigar_score(ops, t, q, match=2, mismatch=-4, q1=4, e1=2, q2=24, e2=1):
    i = j = s = 0
    for o in ops:
        ln, op = int(o) >> 4, int(o) & 0xf
        if op in (0, 7, 8):
            for x in range(ln):
                s += match if t[i + x] == q[j + x] else mismatch
            i += ln
            j += ln
        elif op == 2:
            s -= min(q1 + e1 * ln, q2 + e2 * ln)
            i += ln
        elif op == 1:
            s -= min(q1 + e1 * ln, q2 + e2 * ln)
            j += ln
    return s, i, j


def test_k_cigar_global_path_is_consistent_and_optimal_vs_bruteforce():
    rng = np.random.default_rng(6)
    for _ in range(30):
        t = synth.random_seq(rng, int(rng.integers(5, 120)))
        q = synth.mutate(rng, t, 0.15)
        t, q = t.tobytes().decode(), q.tobytes().decode()
        if not q:
            continue
        ops, r = oracle.k_cigar_ops(t, q, eqx=True)
        s, i, j = cigar_score(ops, t, q)
        assert (i, j) == (len(t), len(q)) and s == r.score
        # eqx ops tell the truth
        ti = qi = 0
        for o in ops:
            ln, op = int(o) >> 4, int(o) & 0xf
            if op == 7:
                assert t[ti:ti + ln] == q[qi:qi + ln]
            if op == 8:
                assert all(a != b for a, b in zip(t[ti:ti + ln], q[qi:qi + ln]))
            if op in (7, 8, 2):
                ti += ln
            if op in (7, 8, 1):
                qi += ln
        # optimality against a plain Gotoh DP with the same dual-affine costs
        assert r.score == gotoh(t, q)


def gotoh(t, q, match=2, mismatch=-4, q1=4, e1=2, q2=24, e2=1):
    NEG = -10 ** 9
    n, m = len(t), len(q)
    H = [[NEG] * (m + 1) for _ in range(n + 1)]
    E1 = [[NEG] * (m + 1) for _ in range(n + 1)]
    F1 = [[NEG] * (m + 1) for _ in range(n + 1)]
    E2 = [[NEG] * (m + 1) for _ in range(n + 1)]
    F2 = [[NEG] * (m + 1) for _ in range(n + 1)]
    H[0][0] = 0
    for i in range(1, n + 1):
        H[i][0] = -min(q1 + e1 * i, q2 + e2 * i)
    for j in range(1, m + 1):
        H[0][j] = -min(q1 + e1 * j, q2 + e2 * j)
    for i in range(1, n + 1):
        for j in range(1, m + 1):
            E1[i][j] = max(H[i - 1][j] - q1, E1[i - 1][j]) - e1
            E2[i][j] = max(H[i - 1][j] - q2, E2[i - 1][j]) - e2
            F1[i][j] = max(H[i][j - 1] - q1, F1[i][j - 1]) - e1
            F2[i][j] = max(H[i][j - 1] - q2, F2[i][j - 1]) - e2
            s = match if t[i - 1] == q[j - 1] else mismatch
            H[i][j] = max(H[i - 1][j - 1] + s, E1[i][j], F1[i][j], E2[i][j], F2[i][j])
    return H[n][m]


def test_k_cigar_extension_stops_at_divergence():
    rng = np.random.default_rng(7)
    t = synth.random_seq(rng, 400).tobytes().decode()
    q = t[:150] + synth.random_seq(rng, 250).tobytes().decode()
    cig, zd, q_e, t_e, _, _ = oracle.k_cigar(t, q, 2, -4, 4, 4, 4, 4, 100, 50)
    assert 145 <= q_e <= 160 and 145 <= t_e <= 160 and zd == 1
    assert oracle.k_cigar(t, "", 2, -4, 4, 4, 4, 4, 100, 50)[2:4] == (0, 0)
