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
