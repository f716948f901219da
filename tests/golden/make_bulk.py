"""Bulk end-to-end fixture from the REFERENCE's own per-read driver (build container only).

    python tests/golden/make_bulk.py            # -> tests/golden/bulk_e2e.json.gz, native_calls.json.gz

For every read of tests/bulk.py's seeded cases the reference's get_readmap_DP_test (mode modules mammap_clrnano /
_ccs / _sensitive) runs over the oracle natives (tests/refrun.py); committed per read: the records (CIGAR as length +
sha1), a status (ok / unmapped / the exception class the reference's worker would have swallowed), the sha1 of the
reference's SAM lines, and BRANCH COUNTERS taken by wrapping the reference's own functions -- second extension pass,
drop_misplaced_alignment_test hits, fix_simple_inv / merge_conjacent_alignment changes, number of guide chains,
heuristic global DP.  For the first reads of every case all calls the reference makes into the absent natives
(`index.map`, `edlib.align`, `k_cigar`) are recorded with their results, so the GPU box can replay them through
the product's `vacmap_index` shim (tests/test_gpu_shim_replay.py).
"""
import gzip
import hashlib
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

N_CALL_READS = 10     # reads per case whose native calls are recorded


def _listify(al):
    return [[tuple(int(v) for v in a) for a in one] for one in al]


def run_shard(task):
    name, shard, n_shards = task
    import bulk
    import oracle
    import refrun
    mode, k, w, over = bulk.CASES[name]
    ref = bulk.reference_for(name)
    reads = bulk.reads_for(name, ref)
    R = refrun.ReferenceRunner(ref, mode=mode, w=w, k=k, **over)
    mod = R.mod
    cnt = {}
    calls = []
    rec_calls = [False]

    # ---- counting wrappers around the reference's own functions (module attributes looked up at call time) ----
    o_extend, o_drop, o_fix, o_merge = mod.extend_func, mod.drop_misplaced_alignment_test, mod.fix_simple_inv, mod.merge_conjacent_alignment
    o_h2w, o_guide = mod.hit2work_1, mod.get_localmap_multi_all_forDP_inv_guide_list

    def extend_func(*a, **kw):
        cnt["extend_calls"] = cnt.get("extend_calls", 0) + 1
        return o_extend(*a, **kw)

    def drop_misplaced(al, iloc, **kw):
        r = o_drop(al, iloc, **kw)
        if r:
            cnt["drop_misplaced"] = cnt.get("drop_misplaced", 0) + 1
        return r

    def fix_simple_inv(al, *a, **kw):
        before = _listify(al)
        r = o_fix(al, *a, **kw)
        if _listify(al) != before:
            cnt["fix_simple_inv"] = cnt.get("fix_simple_inv", 0) + 1
        return r

    def merge_conjacent(al, *a, **kw):
        n0 = len(al)
        r = o_merge(al, *a, **kw)
        if len(al) != n0:
            cnt["merge_conjacent"] = cnt.get("merge_conjacent", 0) + (n0 - len(al))
        return r

    def hit2work_1(one_mapinfo, index2contig, contig2start, testseq_len, skipcost, maxdiff, maxgap, *a, **kw):
        n = len(one_mapinfo)
        fast = n / testseq_len > 5
        if not fast:
            srt = one_mapinfo[oracle.argsort_i64(one_mapinfo[:, 0])]
            g = oracle.chain_global_d_all(srt, R.aligner.k, skipcost[0], maxdiff[0], maxgap)[0]
            fast = g == -1
        if fast:
            cnt["fast_global"] = 1
        cnt["n_anchors"] = n
        return o_h2w(one_mapinfo, index2contig, contig2start, testseq_len, skipcost, maxdiff, maxgap, *a, **kw)

    def guide_list(path_list, *a, **kw):
        cnt["n_chains"] = len(path_list)
        return o_guide(path_list, *a, **kw)

    skip = os.environ.get("BULK_SKIP", "").split(",")
    if "extend" not in skip: mod.extend_func = extend_func
    if "drop" not in skip: mod.drop_misplaced_alignment_test = drop_misplaced
    if "fix" not in skip: mod.fix_simple_inv = fix_simple_inv
    if "merge" not in skip: mod.merge_conjacent_alignment = merge_conjacent
    if "h2w" not in skip: mod.hit2work_1 = hit2work_1
    if "guide" not in skip: mod.get_localmap_multi_all_forDP_inv_guide_list = guide_list

    # ---- recording wrappers around the natives ----
    o_map, o_kc, o_ed = R.aligner.map, mod.mp.k_cigar, mod.edlib.align

    def rec_map(seq, check_num=100, mid_occ=-1):
        out = o_map(seq, check_num=check_num, mid_occ=mid_occ)
        if rec_calls[0]:
            arr = np.array(out, dtype=np.int64).reshape(-1, 4)
            calls.append({"f": "map", "check_num": int(check_num), "mid_occ": int(mid_occ), "n": len(out),
                          "sha": hashlib.sha1(arr.tobytes()).hexdigest()[:16]})
        return out

    def rec_kc(target, query, *a, **kw):
        out = o_kc(target, query, *a, **kw)
        if rec_calls[0]:
            calls.append({"f": "k_cigar", "t": target, "q": query, "a": [int(x) for x in a],
                          "kw": {k_: int(v) for k_, v in kw.items()},
                          "out": [out[0], int(out[1]), int(out[2]), int(out[3]), int(out[4]), int(out[5])]})
        return out

    def rec_ed(query=None, target=None, task="distance", **kw):
        out = o_ed(query=query, target=target, task=task, **kw)
        if rec_calls[0]:
            calls.append({"f": "edlib", "q": query, "t": target, "out": int(out["editDistance"])})
        return out

    class _Ed:
        align = staticmethod(rec_ed)

    class _Mp:
        k_cigar = staticmethod(rec_kc)

        def __getattr__(self, n):
            return getattr(sys.modules["vacmap_index"], n)

    if "map" not in skip: R.aligner.map = rec_map
    if "kc" not in skip: mod.mp = _Mp()
    if "ed" not in skip: mod.edlib = _Ed

    rows = []
    t0 = time.time()
    limit = int(os.environ.get("BULK_LIMIT", "0"))
    for i, (rid, seq) in enumerate(reads):
        if i % n_shards != shard or (limit and i % max(1, len(reads) // limit) != 0 and i >= N_CALL_READS):
            continue
        cnt.clear()
        calls.clear()
        rec_calls[0] = i < N_CALL_READS
        R.last_error = None
        recs = R.align(rid, seq)
        status = "ok" if recs else ("unmapped" if R.last_error is None else type(R.last_error).__name__)
        sam_sha = ""
        squashed = [bulk.squash(r) for r in recs]       # before the emitter: get_bam_dict_str re-orders / rewrites its input list
        if recs:
            o = R.option
            lines = mod.get_bam_dict_str(list(recs), seq.upper(), None, R.contig2iloc, R.contig2seq, o["md"], o["shortcs"],
                                         o["cigar2cg"], o["markunbalancetra"], o)
            sam_sha = hashlib.sha1("\n".join(lines).encode()).hexdigest()[:16]
        c = dict(cnt)
        c["second_pass"] = 1 if c.pop("extend_calls", 0) >= 2 else 0
        rows.append({"i": i, "id": rid, "status": status, "records": squashed, "sam": sam_sha,
                     "cnt": c, "calls": list(calls) if rec_calls[0] else None})
    print(name, "shard", shard, len(rows), "reads", round(time.time() - t0, 1), "s", flush=True)
    return name, rows


def main():
    import bulk
    plan = {"bulk_H": 3, "bulk_H_eqx": 1, "bulk_L_k19": 2, "bulk_S": 2}
    only = sys.argv[1:]
    tasks = [(n, s, ns) for n, ns in plan.items() if (not only or n in only) for s in range(ns)]
    with mp.get_context("spawn").Pool(min(8, len(tasks))) as pool:
        res = pool.map(run_shard, tasks)
    path = os.path.join(HERE, "bulk_e2e.json.gz")
    callpath = os.path.join(HERE, "native_calls.json.gz")
    out = json.load(gzip.open(path, "rt")) if (only and os.path.exists(path)) else {"cases": {}}
    callout = json.load(gzip.open(callpath, "rt")) if (only and os.path.exists(callpath)) else {"cases": {}}
    by = {}
    for name, rows in res:
        by.setdefault(name, []).extend(rows)
    for name, rows in by.items():
        rows.sort(key=lambda r: r["i"])
        mode, k, w, over = bulk.CASES[name]
        callout["cases"][name] = [{"i": r["i"], "calls": r.pop("calls")} for r in rows if r["calls"] is not None]
        for r in rows:
            r.pop("calls", None)
        tot = {}
        for r in rows:
            for kk, v in r["cnt"].items():
                if kk != "n_anchors":
                    tot[kk] = tot.get(kk, 0) + (1 if kk == "n_chains" and v > 1 else v if kk != "n_chains" else 0)
        stat = {}
        for r in rows:
            stat[r["status"]] = stat.get(r["status"], 0) + 1
        out["cases"][name] = {"mode": mode, "k": k, "w": w, "opt": over, "reads": rows, "totals": tot, "status": stat,
                              "max_records": max(len(r["records"]) for r in rows)}
        print(name, len(rows), "reads; status", stat, "; branch totals", tot, "; max records/read", out["cases"][name]["max_records"])
    with gzip.open(path, "wt") as f:
        json.dump(out, f, separators=(",", ":"))
    with gzip.open(callpath, "wt") as f:
        json.dump(callout, f, separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes;", callpath, os.path.getsize(callpath), "bytes")


if __name__ == "__main__" and sys.argv[1:] != ["guide1"]:
    main()


def gen_guide1():
    """Stage golden for local re-seeding: the reference's OWN get_localmap_multi_all_forDP_inv_guide_1
    (mammap_clrnano.py:23069-23345, njit) called on guide chains of bulk reads -> tests/golden/guide1.npz.
    The guide chains are what decode_hit + merge_chain + drop_somechains give for the read (oracle restatements, equal
    to the reference's on every bulk read); they are stored with the anchors the reference function appended for them,
    in its emission order, and the (readstart, readend) it returned."""
    import numba
    import bulk
    import oracle
    import oracle.pipeline as pl
    import refrun
    out = {}
    nj = 0
    for name, picks in (("bulk_H", list(range(0, 12)) + [300, 301, 560, 561, 620, 621, 670, 671]), ("bulk_S", list(range(160, 172)))):
        mode, k, w, over = bulk.CASES[name]
        ref = bulk.reference_for(name)
        reads = bulk.reads_for(name, ref)
        R = refrun.ReferenceRunner(ref, mode=mode, w=w, k=k, **over)
        g1 = R.mod.get_localmap_multi_all_forDP_inv_guide_1

        @numba.njit
        def call(raw, seq, rc, c2s, c2q, skipcost, maxdiff, maxgap):
            one = [(-1, -1, -1, -1)]
            one.pop(0)
            rs, re = g1(-1, -1, one, raw, seq, rc, c2s, c2q, 9, skipcost, maxdiff, maxgap, shift=1)
            res = np.empty((len(one), 4), np.int64)
            for i in range(len(one)):
                res[i, 0] = one[i][0]
                res[i, 1] = one[i][1]
                res[i, 2] = one[i][2]
                res[i, 3] = one[i][3]
            return rs, re, res

        ox = oracle.Index(ref, w=w, k=k)
        opt = R.option
        for i in picks:
            rid, seq = reads[i]
            seq = seq.upper()
            mapq, scores, path_list = pl.decode_hit(ox, seq, len(seq), ox.k, opt, mode)
            if scores == 0.0:
                continue
            rc = pl.revcomp(seq)
            if scores < 0.0:
                seq, rc = rc, seq
            chains = pl.drop_somechains(pl.merge_chain([np.array(p, dtype=np.int64) for p in path_list]))
            for ch in chains:
                rs, re, res = call(np.ascontiguousarray(ch), seq, rc, R.contig2start, R.contig2seq, opt["local_skipcost"],
                                   opt["local_maxdiff"], 99)
                out["j%d_case" % nj] = np.array(name)
                out["j%d_read" % nj] = np.array(i)
                out["j%d_flip" % nj] = np.array(scores < 0.0)
                out["j%d_chain" % nj] = ch.astype(np.int64)
                out["j%d_range" % nj] = np.array([rs, re], np.int64)
                out["j%d_out" % nj] = res
                nj += 1
        print(name, "jobs so far", nj)
    out["n_jobs"] = np.array(nj)
    np.savez_compressed(os.path.join(HERE, "guide1.npz"), **out)
    print("guide1.npz", os.path.getsize(os.path.join(HERE, "guide1.npz")), "bytes,", nj, "jobs,",
          sum(len(out["j%d_out" % j]) for j in range(nj)), "anchors")


if __name__ == "__main__" and sys.argv[1:] == ["guide1"]:
    gen_guide1()
