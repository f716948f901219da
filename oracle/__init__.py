"""ORACLE package -- test infrastructure, NOT product code.

CPU restatement (plain C in ``orc_*.c`` + thin numpy glue here) of the reference's
per-read alignment hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this package.

Parity status:
  * chaining DPs, sorts, local reseed, glue: pinned against the reference's own
    numba functions run in the build container (tests/golden/*.npz).
  * seeding (``vacmap_index.Aligner.map``), ``k_cigar`` and ``edlib`` distance:
    third-party sources absent from /root/reference -> **parity unpinned** for
    the tie-breaking rules; restated from the public minimap2/ksw2/edlib
    algorithms and anchored on the reference's call sites.
"""
import ctypes
import os
import re as _re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

NOPRE = -9999999


def build(force=False):
    out = os.path.join(_HERE, "_build", "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return out


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _declare(_LIB)
    return _LIB


class OrcTables(ctypes.Structure):
    _fields_ = [("extra", ctypes.c_void_p), ("extra_size", ctypes.c_int64),
                ("readgapcost", ctypes.c_void_p), ("log2cache", ctypes.c_void_p),
                ("log2cache_size", ctypes.c_int64)]


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _declare(L):
    i64, dbl, vp, i32 = ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_int
    L.orc_argsort_i64.argtypes = [vp, i64, vp]
    L.orc_argsort_f64.argtypes = [vp, i64, vp]
    L.orc_gapcost_table.argtypes = [i32, i32, i32, vp]
    L.orc_large_readgap_table.argtypes = [i32, i32, vp]
    L.orc_chain_global_d_all.restype = i64
    L.orc_chain_global_d_all.argtypes = [vp, i64, i32, dbl, i64, i64, vp, i64, vp, vp, vp, vp]
    L.orc_chain_linked_d_all.restype = i64
    L.orc_chain_linked_d_all.argtypes = [vp, i64, i64, vp, vp, dbl, i64, i64, i32, dbl, i64, i64, vp, i64, vp, vp, vp, vp, vp]
    L.orc_chain_linked_fast.restype = i64
    L.orc_chain_linked_fast.argtypes = [vp, i64, i64, vp, vp, dbl, i64, i64, i32, dbl, i64, i64, i64, vp, vp, vp, vp]
    L.orc_chain_fast.restype = i64
    L.orc_chain_fast.argtypes = [vp, i64, i32, i32, dbl, i64, i64, i64, vp, vp, vp, vp, vp]
    L.orc_chain_local.restype = i64
    L.orc_chain_local.argtypes = [vp, i64, i32, i32, dbl, i64, i64, vp, vp, vp, vp, vp, vp]
    L.orc_local_traceback.restype = i64
    L.orc_local_traceback.argtypes = [vp, vp, i64, vp]


# ---------------------------------------------------------------------------
# Module-level score tables of the reference (built with numpy exactly like the
# reference builds them at import time, mammap_clrnano.py:15371-15376,
# 26567-26569, 27530), so the values are bit-identical on the same host.
# ---------------------------------------------------------------------------
_TABLES = None


def tables():
    global _TABLES
    if _TABLES is None:
        extra = []
        g = 0
        while True:
            extra.append(min(36, 30 + 0.5 * np.log(max(g, 1)), min(10, g / 100) + min(30, g / 1000)))
            if len(extra) > 1 and extra[-1] == 36:
                break
            g += 1
        extra = np.array(extra, dtype=np.float32)
        readgapcost = np.zeros(100, dtype=np.float32)
        for r in range(1, 100):
            readgapcost[r] = 0.1 * np.log2(r + 1)
        log2cache = np.array([0.5 * np.log2((g + 1)) for g in range(100000)])
        st = OrcTables(_p(extra), len(extra) - 1, _p(readgapcost), _p(log2cache), len(log2cache) - 1)
        _TABLES = dict(extra=extra, readgapcost=readgapcost, log2cache=log2cache, struct=st)
    return _TABLES


def argsort_i64(keys):
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    R = np.empty(len(keys), dtype=np.int64)
    lib().orc_argsort_i64(_p(keys), len(keys), _p(R))
    return R


def argsort_f64(keys):
    keys = np.ascontiguousarray(keys, dtype=np.float64)
    R = np.empty(len(keys), dtype=np.int64)
    lib().orc_argsort_f64(_p(keys), len(keys), _p(R))
    return R


def large_readgap_table(maxgap, large_readgap=30):
    out = np.zeros(maxgap + 1, dtype=np.float32)
    lib().orc_large_readgap_table(maxgap, large_readgap, _p(out))
    return out


def chain_global_d_all(a, kmersize, skipcost, maxdiff, maxgap, max_factor=1000):
    """`_d_all` (mammap_clrnano.py:24828).  a: int64[n,4] sorted by read position.
    Returns (g_max_index | -1, S, P, S_arg, opcount)."""
    a = np.ascontiguousarray(a, dtype=np.int64)
    n = len(a)
    S = np.zeros(n, np.float64)
    P = np.zeros(n, np.int32)
    A = np.zeros(n, np.int32)
    op = np.zeros(1, np.int64)
    t = tables()
    g = lib().orc_chain_global_d_all(_p(a), n, kmersize, float(skipcost), maxdiff, maxgap,
                                     ctypes.addressof(t["struct"]), max_factor, _p(S), _p(P), _p(A), _p(op))
    return g, S, P, A, int(op[0])


_ASM_RG = None


def asm_readgapcost():
    """asm mode's readgapcost_list (mammap_asm.py:16536-16538): float32[100], 0.1 log2(r) -- not clrnano's log2(r + 1)."""
    global _ASM_RG
    if _ASM_RG is None:
        _ASM_RG = np.zeros(100, dtype=np.float32)
        for r in range(1, 100):
            _ASM_RG[r] = 0.1 * np.log2(r)
    return _ASM_RG


def chain_linked_d_all(g_max_scores, g_max_index, pre_S, pre_P, prereadloc, a, kmersize, skipcost, maxdiff, maxgap,
                       max_factor=1000, local=False):
    """asm mode: `linked_..._fine_list_d_all` (mammap_asm.py:21687), or with local=True its second-round twin
    `linked_..._fine_list_all` (:21505).  a: int64[n,4] = carried anchors (len(pre_S) of them) + this batch sorted by
    read position.  Returns (g_max_index | -1, S, P, S_arg, opcount)."""
    a = np.ascontiguousarray(a, dtype=np.int64)
    pre_S = np.ascontiguousarray(pre_S, dtype=np.float64)
    pre_P = np.ascontiguousarray(pre_P, dtype=np.int32)
    n = len(a)
    S = np.zeros(n, np.float64)
    P = np.zeros(n, np.int32)
    A = np.zeros(n, np.int32)
    op = np.zeros(1, np.int64)
    t = tables()
    g = lib().orc_chain_linked_d_all(_p(a), n, len(pre_S), _p(pre_S) if len(pre_S) else None, _p(pre_P) if len(pre_P) else None,
                                     float(g_max_scores), int(g_max_index), int(prereadloc), kmersize, float(skipcost), maxdiff,
                                     maxgap, ctypes.addressof(t["struct"]), max_factor, _p(asm_readgapcost()) if local else None,
                                     _p(S), _p(P), _p(A), _p(op))
    return g, S, P, A, int(op[0])


def chain_linked_fast(g_max_scores, g_max_index, pre_S, pre_P, prereadloc, a, kmersize, skipcost, maxdiff, maxgap, fast_t=5):
    """asm mode: `linked_..._fine_list_d_fast_all` (mammap_asm.py:21872).  Returns (g_max_index, S, P, S_arg_i)."""
    a = np.ascontiguousarray(a, dtype=np.int64)
    pre_S = np.ascontiguousarray(pre_S, dtype=np.float64)
    pre_P = np.ascontiguousarray(pre_P, dtype=np.int32)
    n = len(a)
    S = np.zeros(n, np.float64)
    P = np.zeros(n, np.int32)
    A = np.zeros(n, np.int32)
    t = tables()
    g = lib().orc_chain_linked_fast(_p(a), n, len(pre_S), _p(pre_S) if len(pre_S) else None, _p(pre_P) if len(pre_P) else None,
                                    float(g_max_scores), int(g_max_index), int(prereadloc), kmersize, float(skipcost), maxdiff,
                                    maxgap, fast_t, ctypes.addressof(t["struct"]), _p(S), _p(P), _p(A))
    return g, S, P, A


def chain_fast(a, kmersize, variant, skipcost, maxdiff, maxgap, fast_t=5, large_readgap=30):
    """`_d_fast_all` (variant 0, :25033), `_fine_list_fast` (1, :26938),
    `_fine_list_mismatch_fast` (2, :27891).  Returns (g_max_index, S, P, S_arg_i)."""
    a = np.ascontiguousarray(a, dtype=np.int64)
    n = len(a)
    S = np.zeros(n, np.float64)
    P = np.zeros(n, np.int32)
    A = np.zeros(n, np.int32)
    t = tables()
    rg = None
    if variant == 1:
        rg = t["readgapcost"]
    elif variant == 2:
        rg = large_readgap_table(maxgap, large_readgap)
    g = lib().orc_chain_fast(_p(a), n, kmersize, variant, float(skipcost), maxdiff, maxgap, fast_t,
                             ctypes.addressof(t["struct"]), _p(rg) if rg is not None else None,
                             _p(S), _p(P), _p(A))
    return g, S, P, A


def chain_local(a, kmersize, variant, skipcost, maxdiff, maxgap, large_readgap=30):
    """`_fine_list` (variant 1, :27305) / `_fine_list_mismatch` (variant 2, :28250)
    including the switch to the `_fast` twins.  a sorted by read end.
    Returns (score, path int64[m,4] descending, S, P(int64), used_fast)."""
    a = np.ascontiguousarray(a, dtype=np.int64)
    n = len(a)
    S = np.zeros(n, np.float64)
    P = np.zeros(n, np.int64)
    A = np.zeros(n, np.int64)
    op = np.zeros(1, np.int64)
    t = tables()
    rg = t["readgapcost"] if variant == 1 else large_readgap_table(maxgap, large_readgap)
    g = lib().orc_chain_local(_p(a), n, kmersize, variant, float(skipcost), maxdiff, maxgap,
                              ctypes.addressof(t["struct"]), _p(rg), _p(S), _p(P), _p(A), _p(op))
    used_fast = False
    if g == -2:
        used_fast = True
        g, S, P32, _ = chain_fast(a, kmersize, variant, skipcost, maxdiff, maxgap, 5, large_readgap)
        P = P32.astype(np.int64)
    path = np.zeros((n, 4), np.int64)
    m = lib().orc_local_traceback(_p(a), _p(P), g, _p(path))
    return float(S[g]), path[:m].copy(), S, P, used_fast


# ---------------------------------------------------------------------------
# natives restated from the absent third-party packages (parity unpinned)
# ---------------------------------------------------------------------------
class KcResult(ctypes.Structure):
    _fields_ = [("score", ctypes.c_int32), ("max_t", ctypes.c_int32), ("max_q", ctypes.c_int32),
                ("zdropped", ctypes.c_int32), ("n_cigar", ctypes.c_int32), ("q_e", ctypes.c_int32),
                ("t_e", ctypes.c_int32), ("ndel", ctypes.c_int32), ("nins", ctypes.c_int32)]


_OPS = "MIDNSHP=X"


def _declare_natives(L):
    if getattr(L, "_natives_declared", False):
        return
    i64, vp, i32 = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int32
    L.orc_sketch.restype = i64
    L.orc_sketch.argtypes = [ctypes.c_char_p, i64, i32, i32, vp, vp, i64]
    L.orc_index_build.restype = vp
    L.orc_index_build.argtypes = [ctypes.c_char_p, vp, i32, i32, i32]
    L.orc_index_free.argtypes = [vp]
    L.orc_index_stats.argtypes = [vp, vp, vp, vp]
    L.orc_index_export.argtypes = [vp, vp, vp, vp]
    L.orc_map.restype = i64
    L.orc_map.argtypes = [vp, ctypes.c_char_p, i64, i32, i32, vp, i64]
    L.orc_k_cigar.restype = i32
    L.orc_k_cigar.argtypes = [ctypes.c_char_p, i32, ctypes.c_char_p, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32,
                              vp, i32, ctypes.POINTER(KcResult)]
    L.orc_edit_distance.restype = i64
    L.orc_edit_distance.argtypes = [ctypes.c_char_p, i64, ctypes.c_char_p, i64]
    L._natives_declared = True


def sketch(seq, w, k):
    """minimap2 mm_sketch restated: returns (hash uint64[n], pos_last<<1|strand uint64[n])."""
    L = lib(); _declare_natives(L)
    b = seq.encode() if isinstance(seq, str) else seq
    cap = len(b) + 16
    h = np.zeros(cap, np.uint64)
    y = np.zeros(cap, np.uint64)
    n = L.orc_sketch(b, len(b), w, k, _p(h), _p(y), cap)
    return h[:n].copy(), y[:n].copy()


class Index:
    """Minimizer index over a list of (name, sequence) contigs (global concatenated coordinates)."""

    def __init__(self, contigs, w=10, k=15):
        L = lib(); _declare_natives(L)
        self.w, self.k = w, k
        self.names = [n for n, _ in contigs]
        # mappy's Aligner.seq() hands back the index's 4-bit sequence: every non-ACGT base reads as N
        self.seqs = [_re.sub("[^ACGT]", "N", s.upper().replace("U", "T")) for _, s in contigs]
        self.offsets = np.zeros(len(contigs) + 1, np.int64)
        for i, s in enumerate(self.seqs):
            self.offsets[i + 1] = self.offsets[i] + len(s)
        self._cat = "".join(self.seqs).encode()
        self.h = L.orc_index_build(self._cat, _p(self.offsets), len(contigs), w, k)
        nk, no, mo = np.zeros(1, np.int64), np.zeros(1, np.int64), np.zeros(1, np.int32)
        L.orc_index_stats(self.h, _p(nk), _p(no), _p(mo))
        self.n_keys, self.n_occ, self.mid_occ_default = int(nk[0]), int(no[0]), int(mo[0])

    def export(self):
        keys = np.zeros(self.n_keys, np.uint64)
        start = np.zeros(self.n_keys + 1, np.int64)
        occ = np.zeros(self.n_occ, np.uint64)
        lib().orc_index_export(self.h, _p(keys), _p(start), _p(occ))
        return keys, start, occ

    def map(self, seq, check_num=100, mid_occ=-1):
        """-> int64[n,4] rows (readpos_start, refpos_global_leftmost, strand, len)."""
        L = lib()
        b = seq.encode() if isinstance(seq, str) else seq
        cap = max(4096, 4 * len(b))
        while True:
            rows = np.zeros((cap, 4), np.int64)
            n = L.orc_map(self.h, b, len(b), check_num, mid_occ, _p(rows), cap)
            if n >= 0:
                return rows[:n].copy()
            cap = -n + 16

    def __del__(self):
        try:
            if self.h:
                lib().orc_index_free(self.h)
                self.h = None
        except Exception:
            pass


def k_cigar_ops(target, query, match=2, mismatch=-4, gap_open_1=4, gap_extend_1=2, gap_open_2=24, gap_extend_2=1,
                bw=-1, zdropvalue=-1, eqx=False):
    """-> (ops uint32[n] BAM-encoded, KcResult)."""
    L = lib(); _declare_natives(L)
    t = target.encode() if isinstance(target, str) else target
    q = query.encode() if isinstance(query, str) else query
    cap = len(t) + len(q) + 4
    ops = np.zeros(cap, np.uint32)
    res = KcResult()
    rc = L.orc_k_cigar(t, len(t), q, len(q), match, mismatch, gap_open_1, gap_extend_1, gap_open_2, gap_extend_2,
                       bw, zdropvalue, 1 if eqx else 0, _p(ops), cap, ctypes.byref(res))
    assert rc == 0
    return ops[:res.n_cigar].copy(), res


def ops_to_string(ops):
    return "".join("%d%s" % (int(o) >> 4, _OPS[int(o) & 0xf]) for o in ops)


def k_cigar(target, query, match=2, mismatch=-4, gap_open_1=4, gap_extend_1=2, gap_open_2=24, gap_extend_2=1,
            bw=-1, zdropvalue=-1, eqx=False):
    """vacmap_index.k_cigar contract: (cigar, zdropcode, q_e, t_e, ndel, nins)."""
    ops, r = k_cigar_ops(target, query, match, mismatch, gap_open_1, gap_extend_1, gap_open_2, gap_extend_2,
                         bw, zdropvalue, eqx)
    return ops_to_string(ops), r.zdropped, r.q_e, r.t_e, r.ndel, r.nins


def edit_distance_dp(a, b):
    """plain two-row DP (the definition)"""
    L = lib(); _declare_natives(L)
    a = a.encode() if isinstance(a, str) else a
    b = b.encode() if isinstance(b, str) else b
    return int(L.orc_edit_distance(a, len(a), b, len(b)))


def edit_distance(a, b):
    """Myers bit-vector NW distance (what edlib runs); equals edit_distance_dp (tests/test_oracle_natives.py)."""
    L = lib(); _declare_natives(L)
    if not hasattr(L, "_bv_declared"):
        L.orc_edit_distance_bv.restype = ctypes.c_int64
        L.orc_edit_distance_bv.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_int64]
        L._bv_declared = True
    a = a.encode() if isinstance(a, str) else a
    b = b.encode() if isinstance(b, str) else b
    return int(L.orc_edit_distance_bv(a, len(a), b, len(b)))


def local_reseed_scan(ctg, wins, raw_by_x, seq, rc_seq, k, readstart, readend):
    """orc_local_reseed: the 9-mer table build + read scan + diagonal merge of guide_1 (:23138-23344).
    -> int64[n,4] anchors in the reference's emission order."""
    L = lib(); _declare_natives(L)
    if not hasattr(L, "_reseed_declared"):
        i64, vp, i32 = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int32
        L.orc_local_reseed.restype = i64
        L.orc_local_reseed.argtypes = [ctypes.c_char_p, vp, vp, i32, vp, vp, i64, ctypes.c_char_p, ctypes.c_char_p,
                                       i64, i32, i64, i64, ctypes.POINTER(ctypes.c_void_p)]
        L.orc_free.argtypes = [vp]
        L._reseed_declared = True
    lo = np.array([w[0] for w in wins], dtype=np.int64)
    hi = np.array([w[1] for w in wins], dtype=np.int64)
    gx = np.ascontiguousarray(raw_by_x[:, 0].astype(np.int32))
    gy = np.ascontiguousarray(raw_by_x[:, 1].astype(np.int64))
    out = ctypes.c_void_p()
    n = L.orc_local_reseed(ctg.cat(), _p(lo), _p(hi), len(wins), _p(gx), _p(gy), len(gx), seq.encode(), rc_seq.encode(),
                           len(seq), k, readstart, readend, ctypes.byref(out))
    if n > 0:
        rows = np.ctypeslib.as_array(ctypes.cast(out, ctypes.POINTER(ctypes.c_int64)), shape=(n, 4)).copy()
    else:
        rows = np.zeros((0, 4), np.int64)
    if out.value:
        L.orc_free(out)
    return rows
