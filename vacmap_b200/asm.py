"""asm mode (``-mode asm``), chaining side: host mirror of the batch loops of ``assembly_get_readmap_DP_test``
(``mammap_asm.py:23218-23290`` first round, ``:23306-23404`` second round) over the CUDA linked DPs
(``vm_chain_linked_batch``).  Anchors arrive in batches sorted by read position; after every batch the anchors within
``skipcost + 56`` of the best score are carried into the next one (scores rebased so that the weakest carried anchor
sits at 1000, back-pointers negated = "index into the previous batch"), and the chain is traced back through the
batches at the end.  Same argument meaning and the same quirks as the reference: a batch whose best anchor has no
predecessor is skipped, a carried back-pointer of 0 is followed inside the current batch, and a chain that *starts*
in a carried anchor raises ``IndexError`` (the reference's traceback follows the negated "no predecessor" mark as an
index; its worker then drops the contig).

Still host-side here (numpy): the carry slice and the traceback.  The rest of the contig path (seeding batches,
re-seeding between the rounds, ``ass_extend_func``) follows below the chaining loops.
"""
import time
from contextlib import contextmanager

import numpy as np

# wall seconds spent inside the CUDA entry points, by stage (bench.py --workload cfg4 reads and resets it); everything
# else of a contig's time is the Python host loop
STATS = {}


@contextmanager
def _timed(name):
    t0 = time.perf_counter()
    try:
        yield
    finally:
        STATS[name] = STATS.get(name, 0.0) + time.perf_counter() - t0

from .chain import ChainParams, chain_linked_batch


def _carry(S, P, S_arg, linked, skipcost):
    """:23258-23272 -- the anchors to carry over, in ascending score order."""
    best = S[S_arg[-1]]
    lowest = best - skipcost - 36 - 20
    at = len(S) - 1
    if at <= 0:
        raise Exception("ERROR: ")
    # the reference walks down from the top while the score is above `lowest` (and never tests position 0): the last
    # position whose score is not above it.  No binary search: after a bail-out S_arg is the heuristic DP's order
    # (integer score, then diagonal), which is not monotone in S inside an integer bucket.
    order = S[S_arg]
    below = np.flatnonzero(order <= lowest)
    at = int(below[-1]) if len(below) else 0
    sel = S_arg[at:]
    return S[sel] - order[at] + 1000, (-P[sel]).astype(np.int32), linked[sel]


def linked_chain_path(batches, params=None, second_round=False, ctx=None, dp=None):
    """The chain over all batches as a list of ``(readpos, refpos, strand, len)`` in DESCENDING read order (the
    reference's ``path``), ``[]`` when it has at most one anchor.  ``params``: ChainParams (first round: k 15,
    golbal_skipcost, golbal_maxdiff, maxgap 1000; second round: k 9, local_skipcost, local_maxdiff, maxgap 99).
    ``dp``: replaces the CUDA call (tests of this host logic on a box without a GPU)."""
    params = params or ChainParams()
    if second_round:
        params = ChainParams(params.kmersize, params.skipcost, params.maxdiff, params.maxgap, params.max_factor, params.fast_t,
                             params.large_readgap, 4)
    if dp is None:
        def dp(gs, gi, pS, pP, prl, linked):
            with _timed("linked_dp"):
                r = chain_linked_batch([(gs, gi, pS, pP, prl, linked)], params, ctx=ctx)[0]
            return r.g_max_index, r.S, r.P, r.S_arg
    g_max_scores, g_max_index = 0, 0
    pre_S = np.zeros(0, np.float64)
    pre_P = np.zeros(0, np.int32)
    pre_info = np.zeros((0, 4), np.int64)
    saved = []
    g = None
    for one in batches:
        one = np.asarray(one, dtype=np.int64).reshape(-1, 4)
        if len(one) == 0:
            continue
        if len(pre_info):
            linked = np.concatenate((pre_info, one))
            prereadloc = max(0, int(pre_info[:, 0].max()))
        else:
            linked = one
            prereadloc = int(one[0][0])
        g, S, P, S_arg = dp(g_max_scores, g_max_index, pre_S, pre_P, prereadloc, linked)
        if P[g] < 0:
            continue
        pre_S, pre_P, pre_info = _carry(S, P, S_arg, linked, params.skipcost)
        g_max_index = len(pre_S) - 1
        g_max_scores = pre_S[-1]
        saved.append((linked, P))
    path = []
    for linked, P in reversed(saved):
        # follow the back-pointers on a plain list (a contig's path has ~10^5..10^6 anchors), gather the rows once
        Pl = P.tolist()
        take = int(g)
        idx = [take]
        while Pl[take] >= 0:
            take = Pl[take]                      # IndexError for a chain starting in a carried anchor, as the reference
            idx.append(take)
        if idx[0] >= len(linked):
            raise IndexError("index %d is out of bounds for the batch" % idx[0])
        path.extend(map(tuple, linked[np.array(idx, dtype=np.int64)].tolist()))
        g = abs(int(Pl[take]))
    return path if len(path) > 1 else []


def trim_overlaps(path):
    """:23393-23403 -- ``path`` descending; an anchor reaching into its successor is cut back to the successor's start
    (compared against the untrimmed neighbour); returns the ASCENDING path ``ass_extend_func`` takes."""
    if not len(path):
        return []
    A = np.array(path, dtype=np.int64).reshape(-1, 4)
    pre0 = A[:-1, 0]                              # start of the (untrimmed) neighbour in front
    now = A[1:]
    cut = ~(pre0 >= now[:, 0] + now[:, 3])
    fwd = cut & (now[:, 2] == 1)
    rev = cut & (now[:, 2] != 1)
    out = A.copy()
    new_len = pre0 - now[:, 0]
    out[1:, 3] = np.where(cut, new_len, now[:, 3])
    out[1:, 1] = np.where(rev, now[:, 1] + now[:, 3] - pre0 + now[:, 0], now[:, 1])
    del fwd
    return list(map(tuple, out[::-1].tolist()))


# =================================================================================================================
# The rest of assembly_get_readmap_DP_test (mammap_asm.py:23204-23421) over the CUDA entry points: seeding batches,
# the re-seeding between the rounds, ass_extend_func and the records.  Host logic in Python, like the reference's;
# every hot loop is a kernel call (index.map -> vm_seed_batch_rows, linked DPs -> vm_chain_linked_batch, 9-mer scan ->
# vm_local_reseed_batch, k_cigar -> vm_pairs_batch).
# =================================================================================================================
import bisect
import ctypes
import re

from . import _lib, align
from . import vacmap_index as _vi
from .sam import reverse_complement

_CIG = re.compile(r"(\d+)(\D)")


class ContigDropped(Exception):
    """The reference raises inside this contig and its worker drops it (mammap_asm.py: `except: continue`)."""


def numba_argsort(keys):
    """np.argsort as numba compiles it (the reference sorts inside njit functions, :22754-22755, :22478-)."""
    L = _lib.load()
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    out = np.zeros(len(keys), np.int64)
    L.vm_argsort_i64.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    rc = L.vm_argsort_i64(_lib.ptr(keys), len(keys), _lib.ptr(out))
    if rc != 0:
        raise _lib.VacmapB200Error("vm_argsort_i64 failed (%d)" % rc)
    return out


class Contigs:
    """contig2start / contig2seq / pos2contig of the reference (vacmap:349-367, mammap_asm.py:51-59)."""

    def __init__(self, index):
        self.names = list(index.names)
        self.seqs = [index.seq(n) for n in self.names]
        self.starts = [int(x) for x in index.starts]

    def cid(self, pos):
        return max(bisect.bisect_right(self.starts, pos) - 1, 0)


def link_cigar(c1, c2):
    """:22366-22410 -- two CIGAR strings joined, the boundary ops merged when they are the same op.  Only the boundary is
    looked at (a contig's CIGAR grows to megabytes)."""
    if not c1 or not c2:
        return c1 + c2
    i = len(c1) - 1
    j = i - 1
    while j >= 0 and c1[j].isdigit():
        j -= 1
    k = 0
    while k < len(c2) and c2[k].isdigit():
        k += 1
    if k < len(c2) and c1[i] == c2[k] and j + 1 < i and k > 0:
        return c1[:j + 1] + str(int(c1[j + 1:i]) + int(c2[:k])) + c1[i] + c2[k + 1:]
    return c1 + c2


def link_cigars(pieces):
    """link_cigar folded over the segment CIGARs of a sub-alignment, in linear time: everything in front of the running
    CIGAR's last piece is frozen, only that piece is looked at and rewritten."""
    frozen, last = [], None
    for piece in pieces:
        if last is None:
            last = piece
            continue
        if not last or not piece:
            last = last + piece
            continue
        i = len(last) - 1
        j = i - 1
        while j >= 0 and last[j].isdigit():
            j -= 1
        k = 0
        while k < len(piece) and piece[k].isdigit():
            k += 1
        if k < len(piece) and last[i] == piece[k] and j + 1 < i and k > 0:
            frozen.append(last[:j + 1])
            last = str(int(last[j + 1:i]) + int(piece[:k])) + last[i] + piece[k + 1:]
        else:
            frozen.append(last)
            last = piece
    return "".join(frozen) + (last or "")


def yield_mapinfo(seq, aligner, batch=100000):
    """:22411-22442 -- seeds of 100 kb slices of the contig (`index.map(slice, check_num=-1)`), accumulated until a DP
    batch would exceed 500 000 anchors; every batch sorted by read position with NUMPY's argsort.  Quirk kept: the tail
    flush appends the last slice's anchors a second time.  The slices of a batch are seeded in one launch."""
    slices = [(st, min(st + batch, len(seq))) for st in range(0, len(seq), batch)]
    group = 16
    cache, cache_size, t_st = [], 0, 0
    one = np.zeros((0, 4), np.int64)
    en = 0
    for g0 in range(0, len(slices), group):
        part = slices[g0:g0 + group]
        with _timed("seed"):
            maps = aligner.map_batch([seq[a:b] for a, b in part], check_num=-1, mid_occ=-1, arrays=True)
        for (st, en), rows in zip(part, maps):
            one = np.array(rows, dtype=np.int64).reshape(-1, 4)      # a copy: the read offset is added in place
            if len(one) > 0:
                one[:, 0] += st
            if len(one) + cache_size > 500000:
                if cache_size > 0:
                    if len(one) > 0:
                        cache.append(one)
                    one = np.concatenate(cache)
                    cache_size, cache = 0, []
                yield t_st, en, one[np.argsort(one[:, 0])]
                t_st = en
            else:
                if len(one) > 0:
                    cache.append(one)
                    cache_size += len(one)
    if cache_size > 0:
        if len(one) > 0:
            cache.append(one)
        one = np.concatenate(cache)
        yield t_st, en, one[np.argsort(one[:, 0])]


def guide_windows(raw, ctg, look_span):
    """Reference windows around the guide anchors (:22478- / clrnano:23095-23154): guide reference positions closer than
    max(largest read step + 1000, 5000) share a window, windows reach `look_span` beyond their guides (clipped to the
    contig), a window spanning two contigs makes the clustering start again contig by contig.
    -> ([(lo, hi)] global, guides sorted by read position)."""
    readgap = int(np.abs(np.diff(raw[:, 0])).max()) if len(raw) > 1 else 0
    readgap = max(readgap + 1000, 5000)
    by_y = raw[numba_argsort(raw[:, 1])]

    def cluster(split):
        se = [[int(by_y[0][1]), int(by_y[0][1])]]
        cur = ctg.cid(se[0][0])
        for y in by_y[1:, 1]:
            y = int(y)
            if y - se[-1][1] < readgap and (not split or cur == ctg.cid(y)):
                se[-1][1] = y
            else:
                if se[-1][0] == se[-1][1]:
                    se.pop()
                se.append([y, y])
                cur = ctg.cid(y)
        if se and se[-1][0] == se[-1][1]:
            se.pop()
        return se

    def ranges(se):
        out = []
        for lo_y, hi_y in se:
            c = ctg.cid(lo_y)
            if c != ctg.cid(hi_y):
                return out, True
            cs, n = ctg.starts[c], len(ctg.seqs[c])
            further = min(look_span, lo_y - cs)
            lo, hi, _ = slice(lo_y - further - cs, hi_y + look_span - cs).indices(n)
            out.append((cs + lo, cs + max(hi, lo)))
        return out, False

    wins, retry = ranges(cluster(False))
    if retry:
        wins, _ = ranges(cluster(True))
    return wins, by_y[numba_argsort(by_y[:, 0])]


def collect_second_round_anchors(r_st, r_en, raw, seq, index, ctg, k=9):
    """:22478-22755 -- 9-mers of the read positions [r_st, r_en - k) looked up in the windows around the first-round
    anchors `raw` (2000-base margin), guide-proximity filter and same-diagonal merge as guide_1 (the CUDA re-seeding
    kernels), then sorted by read position with numba's argsort, twice (:22754-22755)."""
    raw = np.ascontiguousarray(raw, dtype=np.int64)
    wins, guides = guide_windows(raw, ctg, 2000)
    with _timed("reseed"):
        rows = align.local_reseed_batch(index, [seq], [(0, wins, guides, int(r_st), int(r_en) - k)])[0]
    if len(rows) == 0:
        return rows
    rows = rows[numba_argsort(rows[:, 0])]
    return rows[numba_argsort(rows[:, 0])]


def yield_second_mapinfo(raw, seq, index, ctg, k=9, batch=100000):
    """:22444-22476 -- batches of second-round anchors along the first-round path `raw` (ascending): a batch ends at a path
    anchor whose successor starts further right, once it reaches `batch` bases past the batch start and holds more than
    300 path anchors; every batch is re-seeded with 20 path anchors of context on both sides."""
    raw = np.ascontiguousarray(raw, dtype=np.int64)
    n = len(raw)
    st_read = st_path = iloc_path = 0
    x, ln = raw[:, 0].tolist(), raw[:, 3].tolist()       # plain ints: the scan below visits every path anchor
    for iloc_path in range(1, n):
        if iloc_path == n - 1 or x[iloc_path + 1] > x[iloc_path]:
            if x[iloc_path] + ln[iloc_path] > st_read + batch and iloc_path - st_path > 300:
                en_read = x[iloc_path]
                rows = collect_second_round_anchors(st_read, en_read, raw[max(0, st_path - 20):min(iloc_path + 20, n)], seq, index, ctg, k)
                if len(rows) > 0:
                    yield rows
                st_path = iloc_path + 1
                st_read = en_read
    if st_read < len(seq):
        rows = collect_second_round_anchors(st_read, len(seq), raw[max(0, st_path - 20):min(iloc_path + 20, n)], seq, index, ctg, k)
        if len(rows) > 0:
            yield rows


# ---- ass_extend_func :23423-23460 ----
def _rebuild_chain_break(ctg, raw, large_cost=50, small_alignment=30):
    """asm's rebuild_chain_break (:13257-): colinear runs of the path (same strand, |readgap - refgap| <= large_cost,
    refgap >= 0, readgap < 100, same contig); singletons and runs spanning < small_alignment read bases are dropped."""
    al, cur = [], [raw[0]]

    def close():
        if len(cur) > 1:
            al.append(list(cur))
        if al and (al[-1][-1][0] + al[-1][-1][3] - al[-1][0][0]) < small_alignment:
            al.pop()

    pre = raw[0]
    for now in raw[1:]:
        ok = False
        if pre[2] == now[2]:
            readgap = now[0] - pre[0] - pre[3]
            refgap = now[1] - pre[1] - pre[3] if pre[2] == 1 else pre[1] - now[1] - now[3]
            ok = abs(readgap - refgap) <= large_cost and refgap >= 0 and readgap < 100 and ctg.cid(pre[1]) == ctg.cid(now[1])
        if ok:
            cur.append(now)
        else:
            close()
            cur = [now]
        pre = now
    # the reference checks the span of the LAST list once more after dropping a final singleton
    if len(cur) > 1:
        al.append(list(cur))
    if not al:
        raise ContigDropped("rebuild_chain_break left nothing")
    if (al[-1][-1][0] + al[-1][-1][3] - al[-1][0][0]) < small_alignment:
        al.pop()
    return al


def _extend_edge(seq, L, al, ctg):
    """extend_edge_test (:2302-2525 of the per-read module, same code in mammap_asm): left then right extension of every
    sub-alignment in order, boundary anchors rewritten to zero-length points; only (q_e, t_e) of k_cigar are used."""
    ext = dict(match=2, mismatch=-4, gap_open_1=4, gap_extend_1=4, gap_open_2=4, gap_extend_2=4, bw=100, zdropvalue=50)
    cap = 20000
    for idx, one in enumerate(al):
        if one[0][0] > 0:
            look = one[0][0] if idx == 0 else one[0][0] - (al[idx - 1][-1][0] + al[idx - 1][-1][3])
            pre = one[0]
            c = ctg.cid(pre[1])
            cs, ref = ctg.starts[c], ctg.seqs[c]
            if pre[2] == 1:
                t_st, q_st = pre[1], pre[0]
                look = min(look, t_st - cs, cap) if min(look, t_st - cs) > cap else min(look, t_st - cs)
                if look != 0:
                    query = seq[max(q_st - look, 0):q_st][::-1]
                    target = ref[slice(t_st - cs - len(query), t_st - cs)][::-1]
                    with _timed("extend"):
                        r = _vi.k_cigar(target, query, **ext)
                    one[0] = (q_st - r[2], t_st - r[3], 1, 0)
            else:
                t_en, q_st = pre[1] + pre[3], pre[0]
                look = min(look, cs + len(ref) - (t_en - 1))
                look = min(look, cap)
                if look != 0:
                    query = seq[max(q_st - look, 0):q_st][::-1]
                    target = reverse_complement(ref[slice(t_en - cs, t_en + len(query) - cs)])[::-1]
                    with _timed("extend"):
                        r = _vi.k_cigar(target, query, **ext)
                    one[0] = (q_st - r[2], t_en + r[3], -1, 0)
        else:
            t = one[0]
            one[0] = (t[0], t[1], 1, 0) if t[2] == 1 else (t[0], t[1] + t[3], -1, 0)
        end = one[-1][0] + one[-1][3]
        if end < L:
            look = L - end if idx + 1 == len(al) else al[idx + 1][0][0] - end
            pre, now = one[-2], one[-1]
            c = ctg.cid(pre[1])
            cs, ref = ctg.starts[c], ctg.seqs[c]
            if pre[2] == 1:
                t_en, q_en = now[1] + now[3], now[0] + now[3]
                look = min(look, cs + len(ref) - (t_en - 1), cap)
                if look != 0:
                    query = seq[q_en:q_en + look]
                    target = ref[slice(t_en - cs, t_en + len(query) - cs)]
                    with _timed("extend"):
                        r = _vi.k_cigar(target, query, **ext)
                    one[-1] = (q_en + r[2], t_en + r[3], 1, 0)
            else:
                t_st, q_en = now[1], now[0] + now[3]
                look = min(look, t_st - cs, cap)
                if look != 0:
                    query = seq[q_en:q_en + look]
                    target = reverse_complement(ref[slice(t_st - cs - len(query), t_st - cs)])
                    with _timed("extend"):
                        r = _vi.k_cigar(target, query, **ext)
                    one[-1] = (q_en + r[2], t_st - r[3], -1, 0)
        else:
            t = one[-1]
            one[-1] = (t[0] + t[3], t[1] + t[3], 1, 0) if t[2] == 1 else (t[0] + t[3], t[1], -1, 0)


def _gaps(pre, now):
    return now[0] - pre[0] - pre[3], (now[1] - pre[1] - pre[3] if pre[2] == 1 else pre[1] - now[1] - now[3])


def _dup_ilocs(al):
    """getdupiloc_numba :16680-16734 (with its `[0][2]` strand-for-length slip)."""
    dup, iloc = set(), 0
    while iloc + 1 < len(al):
        la = al[iloc][-1]
        rp1 = la[0] + la[3]
        ref1, s1 = (la[1] + la[3], 1) if la[2] == 1 else (la[1], -1)
        hit, new_iloc, dupsize, rp2 = False, 0, 0, 0
        for jloc in range(iloc + 1, len(al)):
            if al[jloc][-1][2] == 1:
                ref2, s2 = al[jloc][0][1], 1
            else:
                ref2, s2 = al[jloc][0][1] + al[jloc][0][2], -1
            if s1 != s2:
                continue
            d = ref2 - ref1 if s1 == 1 else ref1 - ref2
            if d < 50:
                hit, new_iloc, dupsize, rp2 = True, jloc, d, al[jloc][0][0]
        if hit:
            readgap = rp2 - rp1
            if iloc + 1 < new_iloc or (dupsize - readgap < -30 and readgap < 30):
                dup.update(range(iloc, new_iloc))
            iloc = new_iloc
        else:
            iloc += 1
    return dup


def _merge_conjacent(al, ctg):
    """merge_conjacent_alignment :16736-16780."""
    if len(al) < 2:
        return
    dup = _dup_ilocs(al)
    iloc = 0
    while iloc + 1 < len(al):
        pre, now = al[iloc][-1], al[iloc + 1][0]
        if iloc in dup or pre[2] != now[2] or ctg.cid(pre[1]) != ctg.cid(now[1]):
            iloc += 1
            continue
        readgap, refgap = _gaps(pre, now)
        if refgap >= 0 and min(readgap, refgap) < 50 and abs(readgap - refgap) < 10000:
            al[iloc].extend(al.pop(iloc + 1))
        else:
            iloc += 1


def _fix_simple_inv(al, ctg, seq):
    """asm's fix_simple_inv (:17159-17198): an inverted middle whose left breakpoint can slide (the bases between are an
    exact match) is re-cut; only the `refen_0 <= refst_1` case exists in this module."""
    for iloc in range(max(len(al) - 2, 0)):
        A, B, C = al[iloc], al[iloc + 1], al[iloc + 2]
        if not (A[0][2] == C[0][2] and A[0][2] != B[0][2] and A[0][2] == 1):
            continue
        c = ctg.cid(A[0][1])
        b0 = ctg.starts[c]
        refen_0, readen_0 = A[-1][1] + A[-1][3] - b0, A[-1][0] + A[-1][3]
        refst_1, readst_1 = B[-1][1] - b0, B[0][0]
        refen_1, readen_1 = B[0][1] + B[0][3] - b0, B[-1][0] + B[-1][3]
        refst_2, readst_2 = C[0][1] - b0, C[0][0]
        if refst_2 - refen_0 != refen_1 - refst_1 or readst_1 - readen_0 + readst_2 - readen_1 != 0:
            continue
        if refst_1 - refen_0 == 0 or refst_1 - refen_0 + refst_2 - refen_1 != 0 or refen_0 > refst_1:
            continue
        if ctg.seqs[c][slice(refen_0, refst_1)] != seq[slice(readen_0, readen_0 - refen_0 + refst_1)]:
            continue
        A[-1] = (readen_0 - refen_0 + refst_1, refst_1 + b0, 1, 0)
        ins = (readen_0 - refen_0 + refst_1, refen_1 + refen_0 - refst_1 + b0, -1, 0)
        while True:
            if not B:
                raise ContigDropped("fix_simple_inv emptied a sub-alignment")
            if ins[0] >= B[0][0]:
                B.pop(0)
            else:
                break
        B.insert(0, ins)


def _query_target(pre, now, seq, rc_seq, L, ctg):
    """get_query_target_for_cigar :5802-5818."""
    if pre[2] == 1:
        c = ctg.cid(pre[1])
        b = ctg.starts[c]
        return ctg.seqs[c][slice(pre[1] - b, now[1] - b)], seq[slice(pre[0], now[0])]
    c = ctg.cid(now[1])
    b = ctg.starts[c]
    return ctg.seqs[c][slice(now[1] + now[3] - b, pre[1] + pre[3] - b)], rc_seq[slice(L - now[0], L - pre[0])]


def _split_alignment(alignment, seq, rc_seq, L, ctg, eqx):
    """asm's split_alignment_test (:22197-22315): anchors are skipped (len < 19 or a gap side < 200) only while both gap
    sides stay below 2000; the segments of a sub-alignment are filled in ONE launch and their CIGARs joined by link_cigar."""
    fwd = alignment[0][2] == 1
    if fwd:
        t = alignment[-1]
        if t[3] != 0:
            alignment[-1] = (t[0] + t[3], t[1] + t[3], 1, 0)
    else:
        t = alignment[0]
        if t[3] != 0:
            alignment[0] = (t[0], t[1] + t[3], -1, 0)
        t = alignment[-1]
        if t[3] != 0:
            alignment[-1] = (t[0] + t[3], t[1], -1, 0)
        alignment = alignment[::-1]
    pre = alignment[0]
    kept, pairs = [pre], []
    for iloc in range(1, len(alignment)):
        now = alignment[iloc]
        readgap = (now[0] - pre[0] - pre[3]) if fwd else (pre[0] - now[0] - now[3])
        refgap = now[1] - pre[1] - pre[3]
        if max(readgap, refgap) < 2000 and (now[3] < 19 or min(readgap, refgap) < 200) and iloc + 1 != len(alignment):
            continue
        target, query = _query_target(pre, now, seq, rc_seq, L, ctg) if fwd else _query_target(now, pre, seq, rc_seq, L, ctg)
        if not (len(target) > 0 and len(query) > 0):
            raise ContigDropped("Failed to compute CIGAR")
        pairs.append((target, query))
        kept.append(now)
        pre = now
    if not pairs:
        raise ContigDropped("no segment to fill")
    with _timed("fill"):
        filled = _vi.k_cigar_batch(pairs, 2, -4, 4, 2, 24, 1, -1, -1, eqx)
    for r in filled:
        if r[0] == "":
            raise ContigDropped("mp.k_cigar ERROR: Failed to compute CIGAR")
    return kept, link_cigars([r[0] for r in filled])


def _records(new_al, cigars, readid, mapq, L, ctg, hardclip):
    """get_onemapinfolist (:20731-20786 of the per-read module; need_reverse is always False in asm mode)."""
    out = []
    clip = "H" if hardclip else "S"
    for al, cg in zip(new_al, cigars):
        c = ctg.cid(al[0][1])
        bias = ctg.starts[c]
        if al[0][2] == 1:
            q_st, q_en = al[0][0], al[-1][0] + al[-1][3]
            r_st, r_en = al[0][1] - bias, al[-1][1] + al[-1][3] - bias
            tail = "%dM" % al[-1][3] if al[-1][3] > 0 else ""
            strand = "+"
        else:
            q_st, q_en = L - al[0][0] - al[0][3], L - al[-1][0]
            r_st, r_en = al[0][1] - bias, al[-1][1] + al[-1][3] - bias
            tail, strand = "", "-"
        full = ("%d%s" % (q_st, clip) if q_st > 0 else "") + cg + tail + ("%d%s" % (L - q_en, clip) if L - q_en > 0 else "")
        qlen = sum(int(n) for n, op in _CIG.findall(full) if op in "MIS=X")
        if qlen != ((q_en - q_st) if hardclip else L):
            raise ContigDropped("cigar length check")
        out.append((readid, ctg.names[c], strand, q_st, q_en, r_st, r_en, mapq, full))
    return out


def ass_extend(path, readid, seq, rc_seq, ctg, opt):
    """ass_extend_func (:23423-23460): no divergence filter, no misplaced-alignment drop; MAPQ 60."""
    L = len(seq)
    al = _rebuild_chain_break(ctg, list(map(tuple, np.asarray(path, dtype=np.int64).reshape(-1, 4).tolist())))
    _extend_edge(seq, L, al, ctg)
    _merge_conjacent(al, ctg)
    _fix_simple_inv(al, ctg, seq)
    new_al, cigars = [], []
    for a in al:
        kept, cg = _split_alignment(a, seq, rc_seq, L, ctg, opt["eqx"])
        new_al.append(kept)
        cigars.append(cg)
    return _records(new_al, cigars, readid, 60, L, ctg, opt["H"])


def assembly_align(readid, seq, index, opt, ctx=None):
    """assembly_get_readmap_DP_test (:23204-23421) for one contig: `onemapinfolist` rows ([] when nothing maps or the
    reference would drop the contig).  Contigs below 500 kb take the per-read path of the module (`-mode S` parameters of
    mammap_asm.py, :23205-23207): here the per-read CUDA pipeline with the asm option set."""
    seq = seq.upper()
    if len(seq) < 500000:
        al = align.Aligner(index, opt, "S")
        return [tuple(r) for r in al.align_batch([(readid, seq)])[0]]
    aligner = _vi.Aligner.__new__(_vi.Aligner)
    aligner._ix, aligner.k, aligner.w = index, index.k, index.w
    ctg = Contigs(index)
    rc_seq = reverse_complement(seq)
    try:
        p1 = ChainParams(kmersize=index.k, skipcost=opt["golbal_skipcost"], maxdiff=opt["golbal_maxdiff"], maxgap=1000)
        path = linked_chain_path((pack[2] for pack in yield_mapinfo(seq, aligner)), p1, ctx=ctx or index.ctx)
        if not path:
            return []
        raw = np.array(path[::-1], dtype=np.int64)
        lk = opt["local_kmersize"]
        p2 = ChainParams(kmersize=lk, skipcost=opt["local_skipcost"], maxdiff=opt["local_maxdiff"], maxgap=99)
        path2 = trim_overlaps(linked_chain_path(yield_second_mapinfo(raw, seq, index, ctg, lk, 100000), p2, second_round=True,
                                                ctx=ctx or index.ctx))
        if not path2:
            return []
        return ass_extend(path2, readid, seq, rc_seq, ctg, opt)
    except (ContigDropped, IndexError, ZeroDivisionError):
        return []
