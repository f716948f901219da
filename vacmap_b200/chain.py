"""Host mirror of the reference's global chaining entry (``hit2work_1`` front half).

``chain_global_batch`` takes, per read, the int64[n,4] anchor array that
``index_object.map()`` returned (``mammap_clrnano.py:23985``) and returns what the
reference computes at ``:23572-23579``: the argsorted anchors, ``S``, ``P``, ``S_arg``
and ``g_max_index`` -- from the CUDA kernels, through the C ABI.
"""
import collections
import ctypes

import numpy as np

from . import _lib

ChainResult = collections.namedtuple("ChainResult", "sorted S P S_arg g_max_index used_fast")


class ChainParams:
    """Arguments of the reference chaining functions (defaults = mode H global)."""

    def __init__(self, kmersize=15, skipcost=40.0, maxdiff=50, maxgap=1000, max_factor=1000, fast_t=5,
                 large_readgap=30, variant=0):
        self.kmersize, self.skipcost, self.maxdiff, self.maxgap = kmersize, skipcost, maxdiff, maxgap
        self.max_factor, self.fast_t, self.large_readgap, self.variant = max_factor, fast_t, large_readgap, variant

    def c(self):
        return _lib.ChainParamsC(self.kmersize, float(self.skipcost), self.maxdiff, self.maxgap, self.max_factor,
                                 self.fast_t, self.large_readgap, self.variant)


def _ragged(anchor_list):
    off = np.zeros(len(anchor_list) + 1, dtype=np.int64)
    for i, a in enumerate(anchor_list):
        off[i + 1] = off[i] + len(a)
    if off[-1] > 0:
        rows = np.ascontiguousarray(np.concatenate([np.asarray(a, dtype=np.int64).reshape(-1, 4) for a in anchor_list]))
    else:
        rows = np.zeros((0, 4), dtype=np.int64)
    return rows, off


class GlobalChainer:
    """Upload / run / download form (device-resident inputs for timing)."""

    def __init__(self, params=None, ctx=None, device=0):
        self.params = params or ChainParams()
        self.ctx = ctx or _lib.default_context(device)
        self.n_reads = 0
        self.total = 0
        self.off = None

    def upload(self, anchor_list, read_lens):
        rows, off = _ragged(anchor_list)
        return self.upload_ragged(rows, off, read_lens)

    def upload_ragged(self, rows, off, read_lens):
        L = _lib.load()
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        off = np.ascontiguousarray(off, dtype=np.int64)
        rl = np.ascontiguousarray(read_lens, dtype=np.int32)
        self.n_reads, self.total, self.off = len(off) - 1, int(off[-1]), off
        p = self.params.c()
        _lib.check(self.ctx.h, L.vm_chain_global_upload(self.ctx.h, ctypes.byref(p), self.n_reads, _lib.ptr(rows),
                                                       _lib.ptr(off), _lib.ptr(rl)))
        return self

    def run(self):
        """Launch the kernels; returns device milliseconds (CUDA events on the ctx stream)."""
        ms = ctypes.c_float(0)
        _lib.check(self.ctx.h, _lib.load().vm_chain_global_run(self.ctx.h, ctypes.byref(ms)))
        return ms.value

    def stage_times(self):
        t = np.zeros(4, dtype=np.float32)
        _lib.check(self.ctx.h, _lib.load().vm_chain_global_times(self.ctx.h, _lib.ptr(t)))
        return dict(pack=float(t[0]), sort=float(t[1]), dp_exact=float(t[2]), dp_fast=float(t[3]))

    def download(self):
        T, n = self.total, self.n_reads
        srt = np.zeros((T, 4), np.int64)
        S = np.zeros(T, np.float64)
        P = np.zeros(T, np.int32)
        A = np.zeros(T, np.int32)
        g = np.zeros(n, np.int64)
        uf = np.zeros(n, np.int32)
        _lib.check(self.ctx.h, _lib.load().vm_chain_global_download(self.ctx.h, _lib.ptr(srt), _lib.ptr(S), _lib.ptr(P),
                                                                   _lib.ptr(A), _lib.ptr(g), _lib.ptr(uf)))
        out = []
        for r in range(n):
            a, b = int(self.off[r]), int(self.off[r + 1])
            out.append(ChainResult(srt[a:b], S[a:b], P[a:b], A[a:b], int(g[r]), bool(uf[r])))
        return out


def chain_global_batch(anchor_list, read_lens, params=None, ctx=None, device=0):
    """One call: host anchors in, per-read ChainResult out (H2D + kernels + D2H)."""
    ch = GlobalChainer(params, ctx, device)
    ch.upload(anchor_list, read_lens)
    ch.run()
    return ch.download()


LocalChainResult = collections.namedtuple("LocalChainResult", "score path used_fast")


def chain_local_batch(anchor_list, read_lens, params, presorted=True, force_fast=False, ctx=None, device=0):
    """Stage-level local chaining (``vm_chain_local_batch``): what ``get_optimal_chain_..._fine_list`` (variant 1),
    ``_fine_list_mismatch`` (variant 2) or their ``_fast`` twins return for each read's int64[n,4] anchors --
    ``(g_max_scores, path)`` -- with the path in ASCENDING read order (the reference's list reversed)."""
    L = _lib.load()
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
    L.vm_chain_local_batch.argtypes = [vp, ctypes.POINTER(_lib.ChainParamsC), i32, i32, i64, vp, vp, vp, vp, vp, vp, vp]
    ctx = ctx or _lib.default_context(device)
    rows, off = _ragged(anchor_list)
    n = len(anchor_list)
    rl = np.ascontiguousarray(read_lens, dtype=np.int32)
    score = np.zeros(n, np.float64)
    path = np.zeros((max(int(off[-1]), 1), 4), np.int64)
    path_off = np.zeros(n + 1, np.int64)
    used_fast = np.zeros(n, np.int32)
    pc = params.c()
    _lib.check(ctx.h, L.vm_chain_local_batch(ctx.h, ctypes.byref(pc), int(presorted), int(force_fast), n, _lib.ptr(rows),
                                             _lib.ptr(off), _lib.ptr(rl), _lib.ptr(score), _lib.ptr(path), _lib.ptr(path_off),
                                             _lib.ptr(used_fast)))
    return [LocalChainResult(float(score[i]), path[path_off[i]:path_off[i + 1]].copy(), int(used_fast[i])) for i in range(n)]


LinkedChainResult = collections.namedtuple("LinkedChainResult", "g_max_index S P S_arg used_fast")


def chain_linked_batch(jobs, params=None, ctx=None, device=0):
    """asm mode (``vm_chain_linked_batch``): the linked global DP with carry-in for a batch of jobs.  Every job is the
    argument list of ``linked_get_optimal_chain_..._fine_list_d_all`` (mammap_asm.py:21687):
    ``(g_max_scores, g_max_index, pre_S, pre_P, prereadloc, one_mapinfo)`` with ``one_mapinfo`` int64[n,4] = the
    ``len(pre_S)`` carried anchors followed by the batch sorted by read position.  Returns what the function returns
    per job, ``(g_max_index, S, P, S_arg)`` -- or, where it bails out on opcount (``used_fast``), what its caller's
    fall-back ``..._d_fast_all`` returns on the same arguments (:23246-23247)."""
    L = _lib.load()
    vp, i64 = ctypes.c_void_p, ctypes.c_int64
    L.vm_chain_linked_batch.argtypes = [vp, ctypes.POINTER(_lib.ChainParamsC), i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    ctx = ctx or _lib.default_context(device)
    params = params or ChainParams()
    rows, off = _ragged([j[5] for j in jobs])
    n = len(jobs)
    total = int(off[-1])
    pre_n = np.array([len(j[2]) for j in jobs], dtype=np.int32)
    head = np.zeros((max(n, 1), 3), np.float64)
    S = np.zeros(max(total, 1), np.float64)
    P = np.zeros(max(total, 1), np.int32)
    for i, (gs, gi, pS, pP, prl, _) in enumerate(jobs):
        head[i] = (float(gs), float(gi), float(prl))
        S[off[i]:off[i] + len(pS)] = pS
        P[off[i]:off[i] + len(pP)] = pP
    A = np.zeros(max(total, 1), np.int32)
    g = np.zeros(max(n, 1), np.int64)
    uf = np.zeros(max(n, 1), np.int32)
    pc = params.c()
    _lib.check(ctx.h, L.vm_chain_linked_batch(ctx.h, ctypes.byref(pc), n, _lib.ptr(rows), _lib.ptr(off), _lib.ptr(pre_n),
                                              _lib.ptr(head), _lib.ptr(S), _lib.ptr(P), _lib.ptr(A), _lib.ptr(g), _lib.ptr(uf)))
    return [LinkedChainResult(int(g[i]), S[off[i]:off[i + 1]].copy(), P[off[i]:off[i + 1]].copy(), A[off[i]:off[i + 1]].copy(),
                              int(uf[i])) for i in range(n)]
