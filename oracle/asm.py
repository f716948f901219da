"""asm mode, first brick: the batch loop around the linked global DP.

TEST INFRASTRUCTURE, like the rest of oracle/.  Restates `assembly_get_readmap_DP_test`'s first round
(mammap_asm.py:23218-23290): anchors arrive in batches sorted by read position; after every batch the anchors within
`skipcost + 56` of the best score are carried over (scores rebased to end at 1000 + ..., back-pointers negated =
"index into the previous batch") and the final chain is traced back through the saved batches.  The DP itself is
`oracle.chain_linked_d_all` (pinned by tests/golden/asm_linked.json.gz); this loop is the build's restatement of
inline reference code and is pinned through the same fixture's `path` entries, produced by running the reference's own
njit function inside a transcription of that loop (tests/golden/make_golden.py::gen_asm_linked).
After an opcount bail-out the heuristic twin `_d_fast_all` takes over (:23246-23247; `oracle.chain_linked_fast`).
"""
import numpy as np

import oracle


def second_round_path(batches, kmersize, skipcost, maxdiff, maxgap=99, dp=None):
    """Second round (mammap_asm.py:23306-23404): the same batch loop over the re-seeded local anchors with
    `linked_..._fine_list_all` (no bail-out, no fall-back), then the overlap trimming of :23393-23402 (an anchor
    reaching into its successor is cut back to the successor's start; the comparison runs against the UNtrimmed
    neighbour) and the reversal to ASCENDING read order that `ass_extend_func` takes.  [] if <= 1 anchors."""
    if dp is None:
        def dp(gs, gi, pS, pP, prl, a):
            g, S, P, A, _ = oracle.chain_linked_d_all(gs, gi, pS, pP, prl, a, kmersize, skipcost, maxdiff, maxgap, local=True)
            return g, S, P, A

    def never(*_):
        raise AssertionError("the second-round DP has no bail-out")
    path = first_round_path(batches, kmersize, skipcost, maxdiff, maxgap, dp=dp, dp_fast=never)
    if not path:
        return []
    pre = path[0]
    for t in range(1, len(path)):
        now = path[t]
        if not pre[0] >= now[0] + now[3]:
            if now[2] == 1:
                path[t] = (now[0], now[1], now[2], pre[0] - now[0])
            else:
                path[t] = (now[0], now[1] + now[3] - pre[0] + now[0], now[2], pre[0] - now[0])
        pre = now
    return path[::-1]


def first_round_path(batches, kmersize, skipcost, maxdiff, maxgap=1000, dp=None, dp_fast=None):
    """batches: iterable of int64[m,4] anchor arrays, each sorted by read position.  Returns the chain as a list of
    (readpos, refpos, strand, len) in DESCENDING read order (as the reference's `path`), [] if it has <= 1 anchors.
    dp / dp_fast: the linked DPs to call (default: the oracle's C restatements) -- the golden generator passes the
    reference's."""
    if dp is None:
        def dp(gs, gi, pS, pP, prl, a):
            g, S, P, A, _ = oracle.chain_linked_d_all(gs, gi, pS, pP, prl, a, kmersize, skipcost, maxdiff, maxgap)
            return g, S, P, A
    if dp_fast is None:
        def dp_fast(gs, gi, pS, pP, prl, a):
            return oracle.chain_linked_fast(gs, gi, pS, pP, prl, a, kmersize, skipcost, maxdiff, maxgap)
    g_max_scores, g_max_index = 0, 0
    pre_S = np.zeros(0, np.float64)
    pre_P = np.zeros(0, np.int32)
    pre_info = np.zeros((0, 4), np.int64)
    saved = []
    pre_g = None
    for one in batches:
        one = np.asarray(one, dtype=np.int64).reshape(-1, 4)
        if len(one) == 0:
            continue
        if len(pre_info) != 0:
            linked = np.concatenate((pre_info, one))
            prereadloc = max(0, int(pre_info[:, 0].max()))           # :23234-23237
        else:
            linked = one
            prereadloc = int(one[0][0])
        pre_g, S, P, S_arg = dp(g_max_scores, g_max_index, pre_S, pre_P, prereadloc, linked)
        if pre_g == -1:                                              # opcount bail-out -> the heuristic twin (:23246-23247)
            pre_g, S, P, S_arg = dp_fast(g_max_scores, g_max_index, pre_S, pre_P, prereadloc, linked)
        if P[pre_g] < 0:
            continue
        g_max_scores = S[S_arg[-1]]
        lowest = g_max_scores - skipcost - 36 - 20
        sliceiloc = len(S) - 1
        if sliceiloc > 0:
            while lowest < S[S_arg[sliceiloc]]:
                sliceiloc -= 1
                if sliceiloc == 0:
                    break
            sliceiloc = max(sliceiloc, 0)
        else:
            raise Exception("ERROR: ")
        sel = S_arg[sliceiloc:]
        pre_S = S[sel] - S[S_arg[sliceiloc]] + 1000
        pre_P = (-P[sel]).astype(np.int32)
        pre_info = linked[sel]
        g_max_index = len(pre_S) - 1
        g_max_scores = pre_S[-1]
        saved.append((linked, P))
    path = []
    g = pre_g
    for linked, P in reversed(saved):
        take = g
        path.append(tuple(int(v) for v in linked[take]))
        while True:
            if P[take] < 0:
                break
            take = P[take]
            path.append(tuple(int(v) for v in linked[take]))
        g = abs(int(P[take]))
    if len(path) <= 1:
        return []
    return path


def collect_second_round_anchors(r_st, r_en, raw, seq, rc_seq, ctg, k=9):
    """Re-seeding between the rounds (mammap_asm.py:22478-22755): 9-mers of the read positions [r_st, r_en - k) looked
    up in the reference windows around the first-round anchors `raw` (int64[m,4]; windows as in guide_1 but with a
    2000-base margin), guide-proximity filter and same-diagonal merge as guide_1, then sorted by read position with
    numba's argsort, twice (:22754-22755).  Built from the pieces of the per-read oracle (oracle/pipeline.py,
    orc_reseed.c): the scan itself is the same code path."""
    import oracle.pipeline as pl
    raw = np.ascontiguousarray(raw, dtype=np.int64)
    wins, raw_x = pl.guide_windows(raw, ctg, look_span=2000)
    rows = oracle.local_reseed_scan(ctg, wins, raw_x, seq, rc_seq, k, int(r_st), int(r_en) - k)
    if len(rows) == 0:
        return rows
    rows = rows[oracle.argsort_i64(rows[:, 0])]
    return rows[oracle.argsort_i64(rows[:, 0])]


def yield_second_mapinfo(raw, seq, rc_seq, ctg, k=9, batch=100000):
    """Batches of second-round anchors along the first-round path `raw` (ascending, int64[m,4]) --
    mammap_asm.py:22444-22476: a batch ends at a path anchor whose successor starts further right, once it reaches
    `batch` bases past the batch start and holds more than 300 path anchors; every batch is re-seeded with 20 path
    anchors of context on both sides."""
    raw = np.ascontiguousarray(raw, dtype=np.int64)
    n = len(raw)
    st_read = st_path = iloc_path = 0
    for now in raw[1:]:
        iloc_path += 1
        if iloc_path == n - 1 or (iloc_path < n - 1 and raw[iloc_path + 1][0] > raw[iloc_path][0]):
            if now[0] + now[3] > st_read + batch and iloc_path - st_path > 300:
                en_read = int(raw[iloc_path][0])
                rows = collect_second_round_anchors(st_read, en_read, raw[max(0, st_path - 20):min(iloc_path + 20, n)], seq, rc_seq, ctg, k)
                if len(rows) > 0:
                    yield rows
                st_path = iloc_path + 1
                st_read = en_read
    if st_read < len(seq):
        rows = collect_second_round_anchors(st_read, len(seq), raw[max(0, st_path - 20):min(iloc_path + 20, n)], seq, rc_seq, ctg, k)
        if len(rows) > 0:
            yield rows


# ---------------------------------------------------------------------------------------------------------------
# End to end for a contig >= 500 kb: assembly_get_readmap_DP_test (mammap_asm.py:23204-23421)
# ---------------------------------------------------------------------------------------------------------------
def link_cigar(c1, c2):
    """:22366-22410 -- the two CIGAR strings joined, their boundary ops merged when equal."""
    import re
    a = re.findall(r"(\d+)(\D)", c1)
    b = re.findall(r"(\d+)(\D)", c2)
    if a and b and a[-1][1] == b[0][1]:
        head = "".join(n + o for n, o in a[:-1])
        tail = "".join(n + o for n, o in b[1:])
        return head + str(int(a[-1][0]) + int(b[0][0])) + a[-1][1] + tail
    return c1 + c2


def yield_mapinfo(seq, index, batch=100000):
    """:22411-22442 -- seeds of 100 kb read slices, accumulated until a DP batch would exceed 500 000 anchors, each
    batch sorted by read position with NUMPY's argsort.  Quirk kept: the tail flush appends the last slice's
    anchors a second time (they are already in the cache)."""
    cache, cache_size, t_st = [], 0, 0
    one = np.zeros((0, 4), np.int64)
    en = 0
    for st in range(0, len(seq), batch):
        en = min(st + batch, len(seq))
        one = np.array(index.map(seq[st:en], -1, -1), dtype=np.int64).reshape(-1, 4)
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


def _rebuild_chain_break_asm(ctg, raw, large_cost, small_alignment):
    """asm's rebuild_chain_break (:13257-): as the per-read one but a sub-alignment only continues over refgap >= 0."""
    import oracle.pipeline as pl
    pre = raw[0]
    al = [[pre]]
    for now in raw[1:]:
        if pre[2] == now[2]:
            readgap = now[0] - pre[0] - pre[3]
            refgap = now[1] - pre[1] - pre[3] if pre[2] == 1 else pre[1] - now[1] - now[3]
            if abs(readgap - refgap) <= large_cost and refgap >= 0 and readgap < 100:
                if ctg.cid(pre[1]) == ctg.cid(now[1]):
                    al[-1].append(now)
                    pre = now
                    continue
        if len(al[-1]) == 1:
            al.pop()
        if len(al) > 0:
            if (al[-1][-1][0] + al[-1][-1][3] - al[-1][0][0]) < small_alignment:
                al.pop()
        al.append([now])
        pre = now
    if len(al[-1]) == 1:
        al.pop()
    if not al:
        raise pl.ReadDropped("rebuild_chain_break: empty")
    if (al[-1][-1][0] + al[-1][-1][3] - al[-1][0][0]) < small_alignment:
        al.pop()
    return al


def _fix_simple_inv_asm(al, ctg, seq):
    """asm's fix_simple_inv (:17159-17198): only the refen_0 < refst_1 case re-cuts the breakpoint."""
    import oracle.pipeline as pl
    if len(al) > 2:
        iloc = 0
        while iloc + 2 < len(al):
            A, B, C = al[iloc], al[iloc + 1], al[iloc + 2]
            if A[0][2] == C[0][2] and A[0][2] != B[0][2] and A[0][2] == 1:
                c = ctg.cid(A[0][1])
                bias0 = ctg.starts[c]
                refen_0 = A[-1][1] + A[-1][3] - bias0
                readen_0 = A[-1][0] + A[-1][3]
                refst_1 = B[-1][1] - bias0
                readst_1 = B[0][0]
                refen_1 = B[0][1] + B[0][3] - bias0
                readen_1 = B[-1][0] + B[-1][3]
                refst_2 = C[0][1] - bias0
                readst_2 = C[0][0]
                if refst_2 - refen_0 == refen_1 - refst_1 and readst_1 - readen_0 + readst_2 - readen_1 == 0:
                    if refst_1 - refen_0 != 0 and refst_1 - refen_0 + refst_2 - refen_1 == 0:
                        if not refen_0 > refst_1:
                            tempref = pl._slice(ctg.seqs[c], refen_0, refst_1)
                            tempquery = pl._slice(seq, readen_0, readen_0 - refen_0 + refst_1)
                            if tempref == tempquery:
                                A[-1] = (readen_0 - refen_0 + refst_1, refst_1 + bias0, 1, 0)
                                ins = (readen_0 - refen_0 + refst_1, refen_1 + refen_0 - refst_1 + bias0, -1, 0)
                                while True:
                                    if not B:
                                        raise pl.ReadDropped("fix_simple_inv emptied a sub-alignment")
                                    if ins[0] >= B[0][0]:
                                        B.pop(0)
                                    else:
                                        break
                                B.insert(0, ins)
            iloc += 1


def _split_alignment_asm(alignment, seq, rc_seq, L, ctg, eqx):
    """asm's split_alignment_test (:22197-22315): anchors are skipped (len < 19 or a gap side < 200) only while both gap
    sides stay below 2000, and the segment CIGARs are joined with link_cigar into ONE string."""
    import oracle.pipeline as pl
    new, cigar = [], None
    fwd = alignment[0][2] == 1
    if fwd:
        if alignment[-1][3] != 0:
            t = alignment[-1]
            alignment[-1] = (t[0] + t[3], t[1] + t[3], 1, 0)
    else:
        if alignment[0][3] != 0:
            t = alignment[0]
            alignment[0] = (t[0], t[1] + t[3], -1, 0)
        if alignment[-1][3] != 0:
            t = alignment[-1]
            alignment[-1] = (t[0] + t[3], t[1], -1, 0)
        alignment = alignment[::-1]
    pre = alignment[0]
    new.append(pre)
    iloc = 1
    while iloc < len(alignment):
        now = alignment[iloc]
        readgap = (now[0] - pre[0] - pre[3]) if fwd else (pre[0] - now[0] - now[3])
        refgap = now[1] - pre[1] - pre[3]
        if max(readgap, refgap) < 2000:
            if now[3] < 19 or min(readgap, refgap) < 200:
                if iloc + 1 != len(alignment):
                    iloc += 1
                    continue
        target, query = pl.query_target(pre, now, seq, rc_seq, L, ctg) if fwd else pl.query_target(now, pre, seq, rc_seq, L, ctg)
        if len(target) > 0 and len(query) > 0:
            cg = oracle.k_cigar(target, query, 2, -4, 4, 2, 24, 1, -1, -1, eqx)[0]
            if cg == "":
                raise pl.ReadDropped("mp.k_cigar ERROR: Failed to compute CIGAR")
            new.append(now)
            cigar = cg if cigar is None else link_cigar(cigar, cg)
        else:
            raise pl.ReadDropped("Failed to compute CIGAR")
        pre = now
        iloc += 1
    if cigar is None:
        raise pl.ReadDropped("cigarlist[-1] == []")
    return new, [cigar]


def ass_extend(path, readid, seq, rc_seq, ctg, opt, kmersize=9):
    """ass_extend_func (:23423-23460): no divergence filter, no misplaced-alignment drop."""
    import oracle.pipeline as pl
    L = len(seq)
    al = _rebuild_chain_break_asm(ctg, [tuple(p) for p in path], 50, 30)
    pl.extend_edge(seq, L, al, ctg)
    pl.merge_conjacent(al, ctg)
    _fix_simple_inv_asm(al, ctg, seq)
    new_al, cigarlist = [], []
    for a in al:
        na, cg = _split_alignment_asm(a, seq, rc_seq, L, ctg, opt["eqx"])
        new_al.append(na)
        cigarlist.append(cg)
    return pl.onemapinfolist(new_al, cigarlist, readid, 60, L, ctg, False, opt["H"])


def assembly_align(readid, seq, index, ctg, opt):
    """assembly_get_readmap_DP_test for a contig of >= 500 000 bases -> onemapinfolist rows ([] when nothing maps).
    Shorter contigs take the asm module's own per-read pipeline (:23205-23207), which is not restated."""
    import oracle.pipeline as pl
    from vacmap_b200.sam import reverse_complement
    if len(seq) < 500000:
        raise NotImplementedError("contigs below 500 kb take mammap_asm's per-read path")
    rc_seq = reverse_complement(seq)
    k = index.k
    batches = (pack[2] for pack in yield_mapinfo(seq, index))
    path = first_round_path(batches, k, opt["golbal_skipcost"], opt["golbal_maxdiff"], 1000)
    if not path:
        return []
    raw = np.array(path[::-1], dtype=np.int64)
    lk = opt["local_kmersize"]
    path2 = second_round_path(yield_second_mapinfo(raw, seq, rc_seq, ctg, lk, 100000), lk, opt["local_skipcost"],
                              opt["local_maxdiff"], 99)
    if not path2:
        return []
    return ass_extend(path2, readid, seq, rc_seq, ctg, opt, lk)
