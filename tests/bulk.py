"""Seeded inputs of the bulk end-to-end fixture (tests/golden/bulk_e2e.json.gz).

The same functions are used by the generator (tests/golden/make_bulk.py, which runs the REFERENCE's own
get_readmap_DP_test in the build container) and by the CPU / GPU parity tests, so only the expected records are
committed, never the sequences.  The read mix is chosen to reach the branches the small e2e fixture does not:
reads the reference leaves unmapped or drops, reads with more than three records, the multi-chain `_mismatch`
local DP, the heuristic `_fast` DPs, `drop_misplaced_alignment_test`, `fix_simple_inv`, the second extension pass,
mode L at -k 19 -w 10, mode S on reads whose donor carries nested SVs drawn from vacsim's grammar
(vacsim/example_parameterfile:1-6: DEL / INS / INV / DUP / TRA / NML, 100-1000 bp, 1-20 adjacent events).
"""
import numpy as np

import synth

_B = np.frombuffer(b"ACGT", dtype=np.uint8)


def _u8(s):
    return np.frombuffer(s.encode(), dtype=np.uint8)


def satellite(rng, unit_len=3000, copies=90, div=0.03):
    """A tandem array: `copies` diverged copies of one random unit (minimizers with ~copies occurrences)."""
    unit = synth.random_seq(rng, unit_len)
    out = []
    for _ in range(copies):
        u = unit.copy()
        n = int(unit_len * div)
        pos = rng.integers(0, unit_len, size=n)
        u[pos] = _B[rng.integers(0, 4, size=n)]
        out.append(u)
    return np.concatenate(out)


def nested_sv(rng, seg, ref, n_events):
    """Apply `n_events` ADJACENT events (vacsim's composite SV: one event per consecutive block) starting at a random
    point of `seg`.  Returns the donor sequence."""
    L = len(seg)
    p = int(rng.integers(500, max(501, L // 2)))
    parts = [seg[:p]]
    for _ in range(n_events):
        sz = int(rng.integers(100, 1000))
        if p + sz >= L - 500:
            break
        blk = seg[p:p + sz]
        kind = ("DEL", "INS", "INV", "DUP", "TRA", "NML")[int(rng.integers(0, 6))]
        if kind == "DEL":
            pass
        elif kind == "INS":
            parts.append(blk)
            parts.append(synth.random_seq(rng, int(rng.integers(100, 1000))))
        elif kind == "INV":
            parts.append(synth._COMP[blk][::-1])
        elif kind == "DUP":
            inv = rng.random() < 0.3
            times = int(rng.integers(2, 5))
            parts.append(blk)
            for _t in range(times - 1):
                parts.append(synth._COMP[blk][::-1] if inv else blk)
        elif kind == "TRA":
            q = int(rng.integers(0, len(ref) - sz))
            other = ref[q:q + sz]
            parts.append(synth._COMP[other][::-1] if rng.random() < 0.5 else other)
        else:
            parts.append(blk)
        p += sz
    parts.append(seg[p:])
    return np.concatenate(parts)


def _finish(rng, seg, err, ratio, rc=None):
    if rc is None:
        rc = rng.random() < 0.5
    if rc:
        seg = synth._COMP[seg][::-1]
    return synth.mutate(rng, seg, err, ratio).tobytes().decode()


def make_mix(contigs, seed, n_plain, n_sv, n_nested, n_chim, n_junk, n_short, n_nrun, n_sat, n_div,
             len_lo=2000, len_hi=12000, err=0.10, ratio=(4, 3, 3), sat_contig=None, n_second=0):
    """-> list of (name, seq).  Names carry the read class."""
    rng = np.random.default_rng(seed)
    arrs = [_u8(s) for _, s in contigs]
    normal = [a for i, a in enumerate(arrs) if contigs[i][0] != sat_contig]
    reads = []

    def draw(ln=None):
        ref = normal[int(rng.integers(0, len(normal)))]
        ln = ln or int(rng.integers(len_lo, len_hi))
        ln = min(ln, len(ref) - 1)
        st = int(rng.integers(0, len(ref) - ln))
        return ref, ref[st:st + ln].copy()

    for i in range(n_plain):
        _, seg = draw()
        reads.append(("plain_%d" % i, _finish(rng, seg, err, ratio)))
    for i in range(n_sv):
        ref, seg = draw()
        reads.append(("sv_%d" % i, _finish(rng, nested_sv(rng, seg, ref, 1), err, ratio)))
    for i in range(n_nested):
        ref, seg = draw(int(rng.integers(max(len_lo, 6000), max(len_hi, 6001))))
        reads.append(("nested_%d" % i, _finish(rng, nested_sv(rng, seg, ref, int(rng.integers(2, 21))), err, ratio)))
    for i in range(n_chim):
        pieces = []
        for _ in range(int(rng.integers(2, 6))):
            _, seg = draw(int(rng.integers(600, 4000)))
            pieces.append(synth._COMP[seg][::-1] if rng.random() < 0.5 else seg)
            if rng.random() < 0.3:
                pieces.append(synth.random_seq(rng, int(rng.integers(50, 600))))
        reads.append(("chim_%d" % i, _finish(rng, np.concatenate(pieces), err, ratio)))
    for i in range(n_junk):
        kind = i % 4
        if kind == 0:
            s = synth.random_seq(rng, int(rng.integers(500, 8000)))
        elif kind == 1:
            s = np.tile(synth.random_seq(rng, int(rng.integers(1, 7))), 3000)[:int(rng.integers(500, 6000))]
        elif kind == 2:
            _, seg = draw(int(rng.integers(200, 900)))        # a short true stretch inside junk
            s = np.concatenate([synth.random_seq(rng, 3000), seg, synth.random_seq(rng, 3000)])
        else:
            _, seg = draw(int(rng.integers(2000, 5000)))
            s = synth.mutate(rng, seg, 0.30)                  # too diverged: seeds, then filtered or dropped
        reads.append(("junk_%d" % i, _finish(rng, s, err, ratio)))
    for i in range(n_short):
        _, seg = draw(int(rng.integers(20, 400)))
        reads.append(("short_%d" % i, _finish(rng, seg, err if i % 2 else 0.0, ratio)))
    for i in range(n_nrun):
        _, seg = draw()
        p = int(rng.integers(0, len(seg) - 400))
        seg[p:p + int(rng.integers(1, 400))] = ord("N")
        reads.append(("nrun_%d" % i, _finish(rng, seg, err, ratio)))
    if sat_contig is not None:
        sat = [a for i, a in enumerate(arrs) if contigs[i][0] == sat_contig][0]
        for i in range(n_sat):
            ln = int(rng.integers(3000, 7000))
            st = int(rng.integers(0, len(sat) - ln))
            reads.append(("sat_%d" % i, _finish(rng, sat[st:st + ln].copy(), 0.01, (1, 1, 1))))
    for i in range(n_div):
        _, seg = draw()
        reads.append(("div_%d" % i, _finish(rng, seg, 0.18 + 0.01 * (i % 6), ratio)))
    # appended last (own generator) so the classes above keep their sequences when this count changes:
    # a short block replaced by same-strand sequence from 2-50 kb away (the "misplaced" middle sub-alignment that
    # drop_misplaced_alignment_test removes, :726-786) plus a deletion and an insertion of similar size elsewhere
    # in the read (pairedindel, :5604-5650) -> the second extension pass with nofilter=True (:24079-24080)
    rng2 = np.random.default_rng(seed + 7919)
    for i in range(n_second):
        ref = normal[int(rng2.integers(0, len(normal)))]
        ln = int(rng2.integers(9000, 12000))
        st = int(rng2.integers(60000, len(ref) - ln - 60000))
        seg = ref[st:st + ln]
        sz = int(rng2.integers(150, 450))
        shift = int(rng2.integers(2000, 50000)) * (1 if rng2.random() < 0.5 else -1)
        p = int(rng2.integers(1500, 3000))
        indel = int(rng2.integers(60, 400))
        d = int(rng2.integers(4500, 5500))
        e = int(rng2.integers(7000, 8000))
        ins = synth.random_seq(rng2, int(indel * rng2.uniform(0.8, 1.0)))
        parts = [seg[:p], ref[st + p + shift:st + p + shift + sz], seg[p + sz:d], seg[d + indel:e], ins, seg[e:]]
        reads.append(("second_%d" % i, _finish(rng2, np.concatenate(parts), err, ratio)))
    return reads


CASES = {
    # name: (mode, k, w, option overrides)
    "bulk_H": ("H", 15, 10, {}),
    "bulk_H_eqx": ("H", 15, 10, {"eqx": True, "md": True}),
    "bulk_L_k19": ("L", 19, 10, {}),
    "bulk_S": ("S", 15, 10, {}),
}


def reference_for(name):
    if name.startswith("bulk_H"):
        ref = synth.make_reference(101, 1500000, n_contigs=3)
        rng = np.random.default_rng(102)
        ref.append(("sat1", satellite(rng).tobytes().decode()))
        return ref
    if name == "bulk_L_k19":
        return synth.make_reference(103, 1500000, n_contigs=2)
    if name == "bulk_S":
        return synth.make_reference(104, 1000000, n_contigs=2)
    raise KeyError(name)


def reads_for(name, ref=None):
    ref = ref or reference_for(name)
    if name == "bulk_H":
        return make_mix(ref, 201, n_plain=300, n_sv=250, n_nested=60, n_chim=60, n_junk=48, n_short=24, n_nrun=20,
                        n_sat=8, n_div=30, sat_contig="sat1", n_second=60)
    if name == "bulk_H_eqx":
        return make_mix(ref, 202, n_plain=60, n_sv=80, n_nested=20, n_chim=20, n_junk=8, n_short=6, n_nrun=6,
                        n_sat=0, n_div=0, sat_contig="sat1")
    if name == "bulk_L_k19":
        return make_mix(ref, 203, n_plain=200, n_sv=200, n_nested=40, n_chim=30, n_junk=16, n_short=8, n_nrun=6,
                        n_sat=0, n_div=0, len_lo=4000, len_hi=15000, err=0.005, ratio=(1, 1, 1), n_second=40)
    if name == "bulk_S":
        return make_mix(ref, 204, n_plain=40, n_sv=120, n_nested=300, n_chim=40, n_junk=12, n_short=6, n_nrun=6,
                        n_sat=0, n_div=12, len_lo=4000, len_hi=10000)
    raise KeyError(name)


def squash(rec):
    """One onemapinfolist row -> compact fixture row: the CIGAR becomes (length, sha1 prefix)."""
    import hashlib
    r = list(rec)
    cg = r[8]
    return [r[1], r[2], int(r[3]), int(r[4]), int(r[5]), int(r[6]), int(r[7]), len(cg), hashlib.sha1(cg.encode()).hexdigest()[:16]]
