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
