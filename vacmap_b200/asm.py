"""asm mode (``-mode asm``), chaining side: host mirror of the batch loops of ``assembly_get_readmap_DP_test``
(``mammap_asm.py:23218-23290`` first round, ``:23306-23404`` second round) over the CUDA linked DPs
(``vm_chain_linked_batch``).  Anchors arrive in batches sorted by read position; after every batch the anchors within
``skipcost + 56`` of the best score are carried into the next one (scores rebased so that the weakest carried anchor
sits at 1000, back-pointers negated = "index into the previous batch"), and the chain is traced back through the
batches at the end.  Same argument meaning and the same quirks as the reference: a batch whose best anchor has no
predecessor is skipped, a carried back-pointer of 0 is followed inside the current batch, and a chain that *starts*
in a carried anchor raises ``IndexError`` (the reference's traceback follows the negated "no predecessor" mark as an
index; its worker then drops the contig).

Still host-side here (numpy): the carry slice and the traceback.  Not mirrored yet: the seeding batches
(``yield_mapinfo``), the re-seeding between the rounds and ``ass_extend_func``.
"""
import numpy as np

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
        take = g
        path.append(tuple(int(v) for v in linked[take]))
        while P[take] >= 0:
            take = P[take]                       # IndexError for a chain starting in a carried anchor, as the reference
            path.append(tuple(int(v) for v in linked[take]))
        g = abs(int(P[take]))
    return path if len(path) > 1 else []


def trim_overlaps(path):
    """:23393-23403 -- ``path`` descending; an anchor reaching into its successor is cut back to the successor's start
    (compared against the untrimmed neighbour); returns the ASCENDING path ``ass_extend_func`` takes."""
    path = list(path)
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
