"""ORACLE per-read pipeline -- test infrastructure, NOT product code.

Pure-Python restatement of the reference's per-read glue around the C restatements in
``orc_*.c``: decode_hit / hit2work_1 bookkeeping, local re-seeding, sub-alignment surgery,
edge extension, fill and record assembly.  Every function cites the reference lines it
follows (``/root/reference/src/vacmap/mammap_clrnano.py`` unless noted).  Pinned against
the reference's own Python run over the same natives (tests/golden/e2e_*.json).
"""
import math

import numpy as np

import oracle

NOPRE = oracle.NOPRE


class ReadDropped(Exception):
    """Mirrors the reference's `except Exception: continue` (:24116-24125): the read emits nothing."""


MODE_CONST = {
    #      accept  max guide chains (None = all)  local maxgap   clamp skipcost in multi-chain local DP
    "H": dict(accept=60.0, max_guides=5, local_maxgap=99, clamp40=False),
    "L": dict(accept=40.0, max_guides=3, local_maxgap=50, clamp40=True),
    "S": dict(accept=40.0, max_guides=None, local_maxgap=99, clamp40=False),
}


def revcomp(s):
    return s.translate(_RC)[::-1]


_RC = str.maketrans("ACGTNacgtn", "TGCANtgcan")


class Contigs:
    """contig2start / contig2seq / pos2contig (:51-59; vacmap:349-367)."""

    def __init__(self, names, seqs):
        self.names = list(names)
        self.seqs = [s.upper() for s in seqs]
        self.starts = []
        off = 0
        for s in self.seqs:
            self.starts.append(off)
            off += len(s)
        self.total = off
        self._cat = None

    def cat(self):
        if self._cat is None:
            self._cat = "".join(self.seqs).encode()
        return self._cat

    def cid(self, pos):
        c = 0
        for i, st in enumerate(self.starts):
            if pos < st:
                break
            c = i
        return c


# ---------------------------------------------------------------------------
# global stage
# ---------------------------------------------------------------------------
def reverse_rough(a, L):
    """get_reversed_chain_numpy_rough :21202-21217"""
    if len(a) < 3:
        return False, a
    nneg = int((a[:, 2] == -1).sum())
    npos = int((a[:, 2] == 1).sum())
    if nneg > npos:
        a = a.copy()
        a[:, 0] = L - a[:, 0] - a[:, 3]
        a[:, 2] *= -1
        return True, np.ascontiguousarray(a[::-1])
    return False, a


def hit2work(a, L, kmersize, skipcost, maxdiff, maxgap, accept, bin_size=100, overlap=0.5):
    """hit2work_1 :23491-23734.  Returns (path_list, mapq, scores_list, secondary_path_list, fast_used)
    with paths as lists of 4-tuples in DESCENDING read order; ([], 0, [], [], _) when rejected."""
    n = len(a)
    fast = n / L > 5
    srt = a[oracle.argsort_i64(a[:, 0])]
    g = -1
    if not fast:
        g, S, P, S_arg, _ = oracle.chain_global_d_all(srt, kmersize, skipcost, maxdiff, maxgap)
    if fast or g == -1:
        fast = True
        g, S, P, S_arg = oracle.chain_fast(srt, kmersize, 0, skipcost, maxdiff, maxgap)
    scores = S[g]
    used = set()
    path_list, scores_list = [], []
    hit = False
    # best chain (:23588-23610)
    path, S_arr = [], []
    take = g
    used.add(take)
    score = S[take]
    while True:
        path.append(tuple(int(v) for v in srt[take]))
        S_arr.append(S[take])
        if P[take] == NOPRE:
            break
        take = int(P[take])
        used.add(take)
    if score > 40:
        hit = True
        scores_list.append(float(score))
        path_list.append(path)
    max_scores = scores if scores > 0 else 0
    # every other chain in descending S (:23617-23640)
    for take in S_arg[::-1]:
        take = int(take)
        if take in used:
            continue
        path = []
        used.add(take)
        score = S[take]
        while True:
            path.append(tuple(int(v) for v in srt[take]))
            if P[take] == NOPRE:
                break
            take = int(P[take])
            if take in used:
                score = score - S[take]
                break
            used.add(take)
        if score > 40:
            scores_list.append(float(score))
            path_list.append(path)
    if not (hit and max_scores > accept):
        return [], 0, [], [], fast
    order = oracle.argsort_f64(np.array(scores_list))[::-1].copy()
    if order[0] != 0:
        for i in range(len(order)):
            if order[i] == 0:
                order[i] = order[0]
                order[0] = 0
                break

    def binset(p):
        return set(x[0] // bin_size for x in p)

    prim_sets = [binset(path_list[order[0]])]
    prim_scores = [[scores_list[order[0]]]]
    for iloc in order[1:]:
        b = binset(path_list[iloc])
        best, pref = 0.0, 0
        for p_loc, ps in enumerate(prim_sets):
            ov = len(ps & b) / min(len(ps), len(b))
            if ov > best:
                best, pref = ov, p_loc
        if best < overlap:
            prim_sets.append(b)
            prim_scores.append([scores_list[iloc]])
        else:
            prim_scores[pref].append(scores_list[iloc])
    m = len(path_list[order[0]])
    f1 = prim_scores[0][0]
    f2 = prim_scores[0][1] if len(prim_scores[0]) >= 2 else 0
    mapq = min(int(40 * (1 - f2 / f1) * min(1, m / 10) * math.log(f1)), 60)   # libm log, as numba lowers np.log
    # select_secondary_alignment :23505-23538
    secondary = []
    if len(path_list) > 1:
        loc2score = np.zeros(L)
        en = L
        for anchor, sc in zip(path_list[0], S_arr):
            st = anchor[0]
            loc2score[st:en] = sc
            en = st
        for iloc in order[1:]:
            one, f2 = path_list[iloc], scores_list[iloc]
            en_loc, st_loc = one[0][0], one[-1][0]
            if en_loc - st_loc < 50:
                continue
            f1 = max(loc2score[en_loc] - loc2score[st_loc], 1)
            if f2 / f1 > 0.9 or abs(f1 - f2) < 40:
                skip = False
                for pri in secondary:
                    pe, ps = pri[0][0], pri[-1][0]
                    ovs = max(min(en_loc, pe) - max(ps, st_loc), 0)
                    if ovs / (en_loc - st_loc) > 0.5:
                        skip = True
                        break
                if not skip:
                    secondary.append(one)
    return path_list, mapq, scores_list, secondary, fast


def _bump(stats, key, by=1):
    if stats is not None:
        stats[key] = stats.get(key, 0) + by


def decode_hit(index, seq, L, kmersize, opt, mode, stats=None):
    """decode_hit :23981-24020 -> (mapq, signed score, return_path_list)"""
    a = index.map(seq, check_num=opt["c"], mid_occ=-1)
    need_reverse, a = reverse_rough(a, L)
    if len(a) <= 2:
        return 0, 0.0, []
    path_list, mapq, scores_list, secondary, fast = hit2work(
        a, L, kmersize, opt["golbal_skipcost"], opt["golbal_maxdiff"], 1000, MODE_CONST[mode]["accept"])
    if fast:
        _bump(stats, "fast_global")
    if len(path_list) == 0:
        return 0, 0.0, []
    ret = [path_list[0]] + list(secondary)
    sc = scores_list[0]
    return mapq, (-sc if need_reverse else sc), ret


# ---------------------------------------------------------------------------
# local stage
# ---------------------------------------------------------------------------
def find_closest(arr, target):
    """findClosest_1 :17560-17581"""
    n = len(arr)
    if target <= arr[0]:
        return arr[0] - target, arr[0] - target, 0, 0
    if target >= arr[n - 1]:
        return target - arr[n - 1], target - arr[n - 1], n - 1, n - 1
    i, j = 0, n
    while i < j:
        mid = (i + j) // 2
        if arr[mid] == target:
            return 0, 0, mid, mid
        if target < arr[mid]:
            j = mid
        else:
            i = mid + 1
    return abs(arr[j - 1] - target), abs(arr[j] - target), j - 1, j


def _windows(raw, readgap, ctg, split_contigs):
    se = [(int(raw[0][1]), int(raw[0][1]))]
    cur = ctg.cid(int(raw[0][1]))
    for item in raw[1:]:
        y = int(item[1])
        if (y - se[-1][1]) < readgap and (not split_contigs or cur == ctg.cid(y)):
            se[-1] = (se[-1][0], y)
        else:
            if se[-1][0] == se[-1][1]:
                se.pop()
            se.append((y, y))
            cur = ctg.cid(y)
    if se[-1][0] == se[-1][1]:
        se.pop()
    return se


def _build_tables(se, ctg, k, look_span):
    single, multi_tab, multi = {}, {}, []
    retry = False
    skip = "N" * k
    for (min_ref, max_ref) in se:
        c = ctg.cid(min_ref)
        if c != ctg.cid(max_ref):
            retry = True
            break
        cs = ctg.starts[c]
        lookfurther = min(look_span, min_ref - cs)
        min_ref -= lookfurther
        max_ref += look_span
        refseq = ctg.seqs[c][min_ref - cs: max_ref - cs]
        for iloc in range(0, len(refseq) - k + 1):
            km = refseq[iloc:iloc + k]
            if km == skip:
                continue
            if km not in single:
                single[km] = min_ref + iloc
            else:
                if km in multi_tab:
                    multi_tab[km].append(min_ref + iloc)
                else:
                    multi_tab[km] = [single[km], min_ref + iloc]
                    multi.append(km)
    for km in multi:
        single.pop(km)
    return single, multi_tab, retry


def guide_windows(raw, ctg, look_span=7000):
    """Window construction of guide_1 (:23095-23154) -> (list of GLOBAL [lo, hi) ranges in insertion
    order, guide chain sorted by read position)."""
    readgap = 0
    pre = raw[0]
    for now in raw[1:]:
        if abs(int(now[0]) - int(pre[0])) > readgap:
            readgap = abs(int(now[0]) - int(pre[0]))
        pre = now
    readgap = max(readgap + 1000, 5000)
    raw = raw[oracle.argsort_i64(raw[:, 1])]

    def ranges(se):
        out = []
        for (min_ref, max_ref) in se:
            c = ctg.cid(min_ref)
            if c != ctg.cid(max_ref):
                return out, True
            cs = ctg.starts[c]
            lookfurther = min(look_span, min_ref - cs)
            lo, hi, _ = slice(min_ref - lookfurther - cs, max_ref + look_span - cs).indices(len(ctg.seqs[c]))
            out.append((cs + lo, cs + max(hi, lo)))
        return out, False

    wins, retry = ranges(_windows(raw, readgap, ctg, False))
    if retry:
        wins, retry = ranges(_windows(raw, readgap, ctg, True))
    raw = raw[oracle.argsort_i64(raw[:, 0])]
    return wins, raw


def local_reseed(out, raw, seq, rc_seq, ctg, k):
    """guide_1 (:23069-23345) with the scan done by the C restatement orc_local_reseed."""
    wins, raw = guide_windows(raw, ctg)
    L = len(seq)
    readstart = max(0, int(raw[0][0]) - 7000)
    readend = min(L - k + 1, int(raw[-1][0]) + 7000)
    rows = oracle.local_reseed_scan(ctg, wins, raw, seq, rc_seq, k, readstart, readend)
    out.extend(tuple(int(v) for v in r) for r in rows)


def local_reseed_py(out, raw, seq, rc_seq, ctg, k):
    """get_localmap_multi_all_forDP_inv_guide_1 :23069-23345, pure Python (cross-check of the C scan).
    `raw`: int64[m,4] guide chain; appends (x, y, strand, len) tuples to `out` in emission order."""
    look_span = 7000
    readgap = 0
    pre = raw[0]
    for now in raw[1:]:
        if abs(int(now[0]) - int(pre[0])) > readgap:
            readgap = abs(int(now[0]) - int(pre[0]))
        pre = now
    readgap = max(readgap + 1000, 5000)
    raw = raw[oracle.argsort_i64(raw[:, 1])]
    se = _windows(raw, readgap, ctg, False)
    single, multi_tab, retry = _build_tables(se, ctg, k, look_span)
    if retry:
        se = _windows(raw, readgap, ctg, True)
        single, multi_tab, retry = _build_tables(se, ctg, k, look_span)
    raw = raw[oracle.argsort_i64(raw[:, 0])]
    L = len(seq)
    readstart = max(0, int(raw[0][0]) - look_span)
    readend = min(L - k + 1, int(raw[-1][0]) + look_span)
    pointdict, point_keys = {}, []
    readposarr = [int(v) for v in raw[:, 0].astype(np.int32)]
    rawx = [int(v) for v in raw[:, 0]]
    rawy = [int(v) for v in raw[:, 1]]

    def hit(iloc, refloc, strand):
        point = refloc - iloc if strand == 1 else -(refloc + iloc)
        if point in pointdict:
            c0, c1, c2, c3 = pointdict[point]
            if (c0 + c3) >= iloc:
                bonus = iloc - (c0 + c3) + k
                if bonus > 0:
                    if c3 + bonus < 20:
                        pointdict[point] = (c0, c1, 1, c3 + bonus) if strand == 1 else (c0, refloc, -1, c3 + bonus)
                    else:
                        out.append((c0, c1, c2, c3))
                        pointdict[point] = (c0 + c3, c1 + c3, 1, bonus) if strand == 1 else (c0 + c3, refloc, -1, bonus)
            else:
                out.append((c0, c1, c2, c3))
                pointdict[point] = (iloc, refloc, strand, k)
        else:
            pointdict[point] = (iloc, refloc, strand, k)
            point_keys.append(point)

    for iloc in range(readstart, readend):
        fwd = seq[iloc:iloc + k]
        rev = rc_seq[L - (iloc + k): L - iloc] if iloc != 0 else ""   # rc[-(0+k):-0] == '' (:23212)
        if fwd == rev:
            continue
        b0, b1, ci0, ci1 = find_closest(readposarr, iloc)
        interval = min(b0 + b1 + 500, 2000)
        r1, r2 = rawy[ci0], rawy[ci1]
        rgap = abs(iloc - rawx[ci0])
        for km, strand in ((fwd, 1), (rev, -1)):
            if km in single:
                locs = (single[km],)
            elif km in multi_tab:
                locs = multi_tab[km]
            else:
                continue
            for refloc in locs:
                diff = abs(rgap - abs(refloc - r1))
                if diff < 500 or (r1 + interval >= refloc >= r1 - interval) or (r2 + interval >= refloc >= r2 - interval):
                    hit(iloc, refloc, strand)
    for key in point_keys:
        out.append(pointdict[key])


def merge_chain(chains):
    """:28529-28569 (chains: list of int64[m,4], each DESCENDING read order)"""
    rest = list(chains[1:])
    if rest:
        order = oracle.argsort_i64(np.array([int(c[-1][0]) for c in rest], dtype=np.int64))
        rest = [rest[i] for i in order]
    iloc = 0
    while iloc < len(rest) - 1:
        jloc = iloc + 1
        while jloc < len(rest):
            a, b = rest[iloc], rest[jloc]
            if int(a[0][0]) + int(a[0][3]) <= int(b[-1][0]) and int(a[0][2]) == int(b[-1][2]):
                readgap = int(b[-1][0]) - int(a[0][0]) - int(a[0][3])
                if int(a[0][2]) == 1:
                    refgap = int(b[-1][1]) - int(a[0][1]) - int(a[0][3])
                else:
                    refgap = int(a[0][1]) - int(b[-1][1]) - int(b[-1][3])
                if abs(readgap - refgap) < 500:
                    rest[iloc] = np.concatenate((b, a))
                    rest.pop(jloc)
                    continue
            jloc += 1
        iloc += 1
    if rest:
        order = oracle.argsort_i64(np.array([len(c) for c in rest], dtype=np.int64))
        rest = [rest[i] for i in order]
    return [chains[0]] + rest


def drop_somechains(chains):
    """:28482-28528"""
    m = len(chains) - 1
    iloclist = [0] * m
    distance = [np.iinfo(np.int64).max] * m
    sc = [[0, 0] for _ in range(m)]
    csc = [[0, 0] for _ in range(m)]
    for item in chains[0]:
        for ci in range(m):
            chain = chains[ci + 1]
            if chain[-1][0] <= item[0] <= chain[0][0]:
                sc[ci][0 if item[2] == 1 else 1] += 1
            while chain[iloclist[ci]][0] > item[0]:
                if iloclist[ci] < len(chain) - 1:
                    iloclist[ci] += 1
                else:
                    break
            t = chain[iloclist[ci]]
            d = abs(int(item[1]) - int(t[1]))
            if d < distance[ci]:
                distance[ci] = d
    for ci in range(m):
        for item in chains[ci + 1]:
            csc[ci][0 if item[2] == 1 else 1] += 1
    out = [chains[0]]
    for ci in range(m):
        keep = (sc[ci][0] > sc[ci][1] and csc[ci][0] > csc[ci][1]) or (sc[ci][0] < sc[ci][1] and csc[ci][0] < csc[ci][1])
        ch = chains[ci + 1]
        if (not keep and distance[ci] < 500) or (int(ch[0][0]) - int(ch[-1][0])) < 100:
            continue
        out.append(ch)
    return out


def local_stage(path_list, seq, rc_seq, ctg, opt, mode, stats=None):
    """get_localmap_multi_all_forDP_inv_guide_list :28479-28589 -> (score, path descending)"""
    mc = MODE_CONST[mode]
    chains = [np.array(p, dtype=np.int64) for p in path_list]
    chains = merge_chain(chains)
    chains = drop_somechains(chains)
    order = oracle.argsort_f64(np.array([1 / len(c) for c in chains]))
    chains = [chains[i] for i in order]
    out = []
    local_reseed(out, chains[0], seq, rc_seq, ctg, 9)
    count = 2
    for ch in chains[1:]:
        local_reseed(out, ch, seq, rc_seq, ctg, 9)
        count += 1
        if mc["max_guides"] is not None and count > mc["max_guides"]:
            break
    a = np.array(out, dtype=np.int64).reshape(-1, 4)
    a = a[oracle.argsort_i64(a[:, 0] + a[:, 3])]
    skip = opt["local_skipcost"]
    if len(chains) > 1:
        _bump(stats, "mismatch_dp")
        if mc["clamp40"]:
            skip = min(skip, 40)
        sc, path, _, _, _ = oracle.chain_local(a, 9, 2, skip, opt["local_maxdiff"], mc["local_maxgap"], 30)
    else:
        sc, path, _, _, _ = oracle.chain_local(a, 9, 1, skip, opt["local_maxdiff"], mc["local_maxgap"])
    return sc, [tuple(int(v) for v in r) for r in path]


# ---------------------------------------------------------------------------
# extension stage
# ---------------------------------------------------------------------------
def rebuild_chain_break(ctg, raw, large_cost, small_alignment=50):
    """:23437-23484 (raw ascending read order)"""
    pre = raw[0]
    al = [[pre]]
    for now in raw[1:]:
        if pre[2] == now[2]:
            readgap = now[0] - pre[0] - pre[3]
            refgap = now[1] - pre[1] - pre[3] if pre[2] == 1 else pre[1] - now[1] - now[3]
            if abs(readgap - refgap) <= large_cost and refgap >= -20 and readgap < 100:
                if ctg.cid(pre[1]) == ctg.cid(now[1]):
                    if refgap >= 0:
                        al[-1].append(now)
                        pre = now
                        continue
                    else:
                        if readgap <= 20:
                            continue
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
        raise ReadDropped("rebuild_chain_break: empty")     # IndexError in the reference
    if (al[-1][-1][0] + al[-1][-1][3] - al[-1][0][0]) < small_alignment:
        al.pop()
    return al


def query_target(pre, now, seq, rc_seq, L, ctg):
    """get_query_target_for_cigar :5802-5818"""
    if pre[2] == 1:
        c = ctg.cid(pre[1])
        b = ctg.starts[c]
        return _slice(ctg.seqs[c], pre[1] - b, now[1] - b), _slice(seq, pre[0], now[0])
    c = ctg.cid(now[1])
    b = ctg.starts[c]
    return _slice(ctg.seqs[c], now[1] + now[3] - b, pre[1] + pre[3] - b), _slice(rc_seq, L - now[0], L - pre[0])


def _slice(s, a, b):
    """Python slice semantics incl. negative indices, as the reference's str slicing"""
    return s[a:b]


def extend_edge(seq, L, al, ctg, san=1):
    """extend_edge_test :2302-2525.  Only (q_e, t_e) of the z-drop extension are used."""
    max_extend = 20000
    for idx in range(len(al)):
        one = al[idx]
        if one[0][0] > 0:
            pre_idx = max(idx - san, 0)
            if idx == 0 or idx - san < 0:
                looksize = one[0][0]
            else:
                looksize = one[0][0] - (al[pre_idx][-1][0] + al[pre_idx][-1][3])
            pre = one[0]
            c = ctg.cid(pre[1])
            cs, clen = ctg.starts[c], len(ctg.seqs[c])
            if pre[2] == 1:
                target_st, query_st = pre[1], pre[0]
                looksize = min(looksize, target_st - cs)
                if looksize > max_extend:
                    looksize = max_extend
                if looksize != 0:
                    query = _slice(seq, max(query_st - looksize, 0), query_st)[::-1]
                    target = _slice(ctg.seqs[c], target_st - cs - len(query), target_st - cs)[::-1]
                    _, _, q_e, t_e, _, _ = oracle.k_cigar(target, query, 2, -4, 4, 4, 4, 4, 100, 50)
                    one[0] = (query_st - q_e, target_st - t_e, 1, 0)
            else:
                target_en, query_st = pre[1] + pre[3], pre[0]
                looksize = min(looksize, cs + clen - (target_en - 1))
                if looksize > max_extend:
                    looksize = max_extend
                if looksize != 0:
                    query = _slice(seq, max(query_st - looksize, 0), query_st)[::-1]
                    target = revcomp(_slice(ctg.seqs[c], target_en - cs, target_en + len(query) - cs))[::-1]
                    _, _, q_e, t_e, _, _ = oracle.k_cigar(target, query, 2, -4, 4, 4, 4, 4, 100, 50)
                    one[0] = (query_st - q_e, target_en + t_e, -1, 0)
        else:
            t = one[0]
            one[0] = (t[0], t[1], 1, 0) if t[2] == 1 else (t[0], t[1] + t[3], -1, 0)
        if (one[-1][0] + one[-1][3]) < len(seq):
            nxt = min(idx + san, len(al))
            if nxt == len(al):
                looksize = L - (one[-1][0] + one[-1][3])
            else:
                looksize = al[nxt][0][0] - (one[-1][0] + one[-1][3])
            pre, now = one[-2], one[-1]
            c = ctg.cid(pre[1])
            cs, clen = ctg.starts[c], len(ctg.seqs[c])
            if pre[2] == 1:
                target_en, query_en = now[1] + now[3], now[0] + now[3]
                looksize = min(looksize, cs + clen - (target_en - 1))
                if looksize > max_extend:
                    looksize = max_extend
                if looksize != 0:
                    query = _slice(seq, query_en, query_en + looksize)
                    target = _slice(ctg.seqs[c], target_en - cs, target_en + len(query) - cs)
                    _, _, q_e, t_e, _, _ = oracle.k_cigar(target, query, 2, -4, 4, 4, 4, 4, 100, 50)
                    one[-1] = (query_en + q_e, target_en + t_e, 1, 0)
            else:
                target_st, query_en = now[1], now[0] + now[3]
                looksize = min(looksize, target_st - cs)
                if looksize > max_extend:
                    looksize = max_extend
                if looksize != 0:
                    query = _slice(seq, query_en, query_en + looksize)
                    target = revcomp(_slice(ctg.seqs[c], target_st - cs - len(query), target_st - cs))
                    _, _, q_e, t_e, _, _ = oracle.k_cigar(target, query, 2, -4, 4, 4, 4, 4, 100, 50)
                    one[-1] = (query_en + q_e, target_st - t_e, -1, 0)
        else:
            t = one[-1]
            one[-1] = (t[0] + t[3], t[1] + t[3], 1, 0) if t[2] == 1 else (t[0] + t[3], t[1], -1, 0)


def _gaps(pre, now):
    readgap = now[0] - pre[0] - pre[3]
    refgap = now[1] - pre[1] - pre[3] if pre[2] == 1 else pre[1] - now[1] - now[3]
    return readgap, refgap


def drop_misplaced(al, iloc):
    """drop_misplaced_alignment_test :726-786"""
    a, b, c = al[iloc], al[iloc + 1], al[iloc + 2]
    if a[0][2] == b[0][2] and a[0][2] == c[0][2]:
        mid = b[-1][0] + b[-1][3] - b[0][0]
        if mid > 1000:
            return False
        readgap, refgap = _gaps(a[-1], b[0])
        if abs(refgap) < 100000:
            DEL = INS = 0
            if readgap - refgap < -30:
                DEL += 1
            elif readgap - refgap > 30:
                INS += 1
            else:
                return False
            gap_1 = abs(readgap - refgap)
            readgap, refgap = _gaps(b[-1], c[0])
            if abs(refgap) < 100000:
                if readgap - refgap < -30:
                    DEL += 1
                elif readgap - refgap > 30:
                    INS += 1
                else:
                    return False
                gap_2 = abs(readgap - refgap)
                if DEL == 1 and INS == 1 and (mid < 500 or max(gap_1, gap_2) / mid > 0.5):
                    al.pop(iloc + 1)
                    return True
    return False


def getdupiloc(al):
    """getdupiloc_numba :16680-16734 (incl. the `[0][2]` strand-for-length quirk, Appendix A5)"""
    dup = []
    if len(al) >= 2:
        iloc = 0
        while iloc + 1 < len(al):
            readpos_1 = al[iloc][-1][0] + al[iloc][-1][3]
            if al[iloc][-1][2] == 1:
                refpos_1, strand_1 = al[iloc][-1][1] + al[iloc][-1][3], 1
            else:
                refpos_1, strand_1 = al[iloc][-1][1], -1
            jloc, hit, dupsize = iloc, False, 0
            new_iloc = readpos_2 = 0
            while jloc + 1 < len(al):
                jloc += 1
                if al[jloc][-1][2] == 1:
                    refpos_2, strand_2 = al[jloc][0][1], 1
                else:
                    refpos_2, strand_2 = al[jloc][0][1] + al[jloc][0][2], -1
                if strand_1 != strand_2:
                    continue
                d = refpos_2 - refpos_1 if strand_1 == 1 else refpos_1 - refpos_2
                if d < 50:
                    new_iloc, dupsize, readpos_2, hit = jloc, d, al[jloc][0][0], True
            if hit:
                readgap = readpos_2 - readpos_1
                if (iloc + 1) < new_iloc or ((dupsize - readgap) < -30 and readgap < 30):
                    dup.extend(range(iloc, new_iloc))
                iloc = new_iloc
            else:
                iloc += 1
    return dup


def merge_conjacent(al, ctg):
    """merge_conjacent_alignment :16736-16780"""
    if len(al) >= 2:
        iloc = 0
        dup = getdupiloc(al)
        while iloc + 1 < len(al):
            if iloc in dup:
                iloc += 1
                continue
            pre, now = al[iloc][-1], al[iloc + 1][0]
            if pre[2] != now[2] or ctg.cid(pre[1]) != ctg.cid(now[1]):
                iloc += 1
                continue
            readgap, refgap = _gaps(pre, now)
            if refgap < 0:
                iloc += 1
                continue
            if min(readgap, refgap) < 50 and abs(readgap - refgap) < 10000:
                al[iloc] = al[iloc] + al[iloc + 1]
                al.pop(iloc + 1)
            else:
                iloc += 1


def _rc_n(s):
    return "".join({"A": "T", "T": "A", "G": "C", "C": "G"}.get(c, "N") for c in s[::-1])


def fix_simple_inv(al, ctg, seq):
    """fix_simple_inv :24226-24312"""
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
                        if refen_0 > refst_1:
                            tempref = _rc_n(_slice(ctg.seqs[c], refen_1, refen_1 + refen_0 - refst_1))
                            tempquery = _slice(seq, readen_0 - refen_0 + refst_1, readen_0)
                            if tempref == tempquery:
                                bias = refen_0 - refst_1
                                C[0] = (readst_2 - bias, refst_2 - bias + bias0, 1, 0)
                                ins = (readst_2 - bias, refen_0 + bias0, -1, 0)
                                while True:
                                    if not B:
                                        raise ReadDropped("fix_simple_inv emptied a sub-alignment")
                                    if ins[0] <= B[-1][0] + B[-1][3]:
                                        B.pop()
                                    else:
                                        break
                                B.append(ins)
                        else:
                            tempref = _slice(ctg.seqs[c], refen_0, refst_1)
                            tempquery = _slice(seq, readen_0, readen_0 - refen_0 + refst_1)
                            if tempref == tempquery:
                                A[-1] = (readen_0 - refen_0 + refst_1, refst_1 + bias0, 1, 0)
                                ins = (readen_0 - refen_0 + refst_1, refen_1 + refen_0 - refst_1 + bias0, -1, 0)
                                while True:
                                    if not B:
                                        raise ReadDropped("fix_simple_inv emptied a sub-alignment")
                                    if ins[0] >= B[0][0]:
                                        B.pop(0)
                                    else:
                                        break
                                B.insert(0, ins)
            iloc += 1


def split_alignment(alignment, seq, rc_seq, L, ctg, eqx):
    """split_alignment_test :21505-21617 -> (new_alignment (one list), list of per-segment cigar strings)"""
    new, cigars = [], []
    if alignment[0][2] == 1:
        if alignment[-1][3] != 0:
            t = alignment[-1]
            alignment[-1] = (t[0] + t[3], t[1] + t[3], 1, 0)
        pre = alignment[0]
        new.append(pre)
        iloc = 1
        while iloc < len(alignment):
            now = alignment[iloc]
            readgap = now[0] - pre[0] - pre[3]
            refgap = now[1] - pre[1] - pre[3]
            if now[3] < 19 or min(readgap, refgap) < 200:
                if iloc + 1 != len(alignment):
                    iloc += 1
                    continue
            target, query = query_target(pre, now, seq, rc_seq, L, ctg)
            if len(target) > 0 and len(query) > 0:
                cigars.append(oracle.k_cigar(target, query, 2, -4, 4, 2, 24, 1, -1, -1, eqx)[0])
                new.append(now)
            else:
                raise ReadDropped("Failed to compute CIGAR")
            pre = now
            iloc += 1
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
            readgap = pre[0] - now[0] - now[3]
            refgap = now[1] - pre[1] - pre[3]
            if now[3] < 19 or min(readgap, refgap) < 200:
                if iloc + 1 != len(alignment):
                    iloc += 1
                    continue
            target, query = query_target(now, pre, seq, rc_seq, L, ctg)
            if len(target) > 0 and len(query) > 0:
                cigars.append(oracle.k_cigar(target, query, 2, -4, 4, 2, 24, 1, -1, -1, eqx)[0])
                new.append(now)
            else:
                raise ReadDropped("Failed to compute CIGAR")
            pre = now
            iloc += 1
    if not cigars:
        raise ReadDropped("cigarlist[-1] == []")
    return new, cigars


def cigar_query_len(c):
    n = num = 0
    for ch in c:
        if ch.isdigit():
            num = num * 10 + ord(ch) - 48
        else:
            if ch in "MIS=X":
                n += num
            num = 0
    return n


def onemapinfolist(new_al, cigarlist, readid, mapq, L, ctg, need_reverse, hardclip):
    """get_onemapinfolist :20731-20838"""
    clip = "H" if hardclip else "S"
    out = []
    for iloc, al in enumerate(new_al):
        c = ctg.cid(al[0][1])
        bias = ctg.starts[c]
        cig = "".join(cigarlist[iloc])
        if al[0][2] == 1:
            q_st, q_en = al[0][0], al[-1][0] + al[-1][3]
            t_st, t_en = al[0][1], al[-1][1] + al[-1][3]
            top = (str(q_st) + clip) if q_st > 0 else ""
            tail = (str(L - q_en) + clip) if (L - q_en) > 0 else ""
            if al[-1][3] > 0:
                tail = str(int(al[-1][3])) + "M" + tail
            strand = "-" if need_reverse else "+"
        else:
            q_st, q_en = L - al[0][0] - al[0][3], L - al[-1][0]
            t_st, t_en = al[0][1], al[-1][1] + al[-1][3]
            top = (str(q_st) + clip) if q_st > 0 else ""
            tail = (str(L - q_en) + clip) if (L - q_en) > 0 else ""
            strand = "+" if need_reverse else "-"
        out.append((readid, ctg.names[c], strand, q_st, q_en, t_st - bias, t_en - bias, mapq, top + cig + tail))
    for line in out:
        want = (line[4] - line[3]) if hardclip else L
        if want != cigar_query_len(line[-1]):
            raise ReadDropped("cigar length check")
    return out[::-1] if need_reverse else out


def paired_indel(cigars, indelsize=30):
    """pairedindel :5604-5650"""
    indel = []
    for c in cigars:
        num = 0.0
        for ch in c:
            if ch.isdigit():
                num = num * 10.0 + (ord(ch) - 48)
            else:
                if ch in "ID" and num > indelsize:
                    indel.append(num)
                num = 0.0
    indel.sort()
    pre = 0
    for now in indel:
        if max(pre, now) > 0 and (min(pre, now) / max(pre, now)) > 0.7:
            return True
        pre = now
    return False


def extend_func(raw, readid, mapq, seq, rc_seq, L, ctg, need_reverse, opt, nofilter, stats=None):
    """extend_func :19238-19303 -> (onemapinfolist, filtered)"""
    al = rebuild_chain_break(ctg, raw, opt["local_maxdiff"])
    i = 0
    while i < len(al):
        target, query = query_target(al[i][0], al[i][-1], seq, rc_seq, L, ctg)
        m = min(len(target), len(query))
        if m == 0:
            raise ReadDropped("division by zero in divergence filter")
        if oracle.edit_distance(query, target) / m > opt["maxdivergence"]:
            al.pop(i)
        else:
            i += 1
    extend_edge(seq, L, al, ctg)
    n0 = len(al)
    filtered = False
    if len(al) > 2 and not nofilter:
        iloc = 0
        while iloc < len(al) - 2:
            if not drop_misplaced(al, iloc):
                iloc += 1
            else:
                _bump(stats, "drop_misplaced")
    if len(al) < n0:
        filtered = True
        extend_edge(seq, L, al, ctg)
    n1 = len(al)
    merge_conjacent(al, ctg)
    if len(al) != n1:
        _bump(stats, "merge_conjacent", n1 - len(al))
    before = [list(x) for x in al]
    fix_simple_inv(al, ctg, seq)
    if [list(x) for x in al] != before:
        _bump(stats, "fix_simple_inv")
    new_al, cigarlist = [], []
    for a in al:
        na, cg = split_alignment(a, seq, rc_seq, L, ctg, opt["eqx"])
        new_al.append(na)
        cigarlist.append(cg)
    return onemapinfolist(new_al, cigarlist, readid, mapq, L, ctg, need_reverse, opt["H"]), filtered


def align_read(readid, seq, index, ctg, opt, mode="H", stats=None):
    """get_readmap_DP_test :24023-24084 -> list of 9-tuples (possibly empty)"""
    seq = seq.upper()
    L = len(seq)
    rc_seq = revcomp(seq)
    try:
        mapq, scores, path_list = decode_hit(index, seq, L, index.k, opt, mode, stats)
        if scores == 0.0:
            return []
        need_reverse = scores < 0.0
        if need_reverse:
            seq, rc_seq = rc_seq, seq
        sc, raw = local_stage(path_list, seq, rc_seq, ctg, opt, mode, stats)
        if len(raw) <= 1:
            return []
        asc = raw[::-1]
        recs, filtered = extend_func(list(asc), readid, mapq, seq, rc_seq, L, ctg, need_reverse, opt, opt["nodiscard"], stats)
        if len(recs) == 0:
            return []
        if (not opt["nodiscard"]) and filtered and paired_indel([r[-1] for r in recs]):
            _bump(stats, "second_pass")
            recs, filtered = extend_func(list(asc), readid, mapq, seq, rc_seq, L, ctg, need_reverse, opt, True, stats)
        return recs
    except (ReadDropped, ZeroDivisionError, IndexError):
        _bump(stats, "dropped")
        return []
