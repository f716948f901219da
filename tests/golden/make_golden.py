"""Generate golden vectors by running the REFERENCE's own numba functions.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py [chain]
Outputs small .npz fixtures under tests/golden/.  The GPU box has no reference
tree; tests there compare against these files and against the C oracle.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import refimport  # noqa: E402
import synth  # noqa: E402

P = "get_optimal_chain_sortbyreadpos_forSV_inv_test_merged_fine_list"


def gen_chain():
    import numba
    m = refimport.load_mode("clrnano")

    @numba.njit
    def nb_argsort(a):
        return np.argsort(a)

    rng = np.random.default_rng(20261017)
    out = {}
    # --- argsort permutation cases (numba quicksort) ---
    ncase = 0
    for n in (1, 2, 3, 14, 15, 16, 17, 40, 100, 333, 1000, 2500):
        for div in (1, 3, 50):
            a = rng.integers(0, max(2, n // div), size=n).astype(np.int64)
            out["sort_%d_in" % ncase] = a
            out["sort_%d_out" % ncase] = nb_argsort(a).astype(np.int32)
            ncase += 1
    out["sort_count"] = np.array(ncase)
    # --- global DP cases ---
    dall = getattr(m, P + "_d_all")
    dfast = getattr(m, P + "_d_fast_all")
    cases = []
    for t in range(14):
        if t % 3 == 0:
            a = synth.anchors_tieheavy(rng, n=int(rng.integers(3, 500)))
        else:
            a = synth.anchors_global(rng, n_true=int(rng.integers(10, 500)), n_noise=int(rng.integers(0, 900)))
        cases.append(a)
    # noise only, large: triggers the opcount bail-out (:24914) -> g = -1
    cases.append(synth.anchors_global(rng, n_true=0, n_noise=3200, repeats=False))
    for ci, a in enumerate(cases):
        out["g_%d_raw" % ci] = a.astype(np.int32)   # as index.map() would hand them over (unsorted)
        a = a[nb_argsort(a[:, 0])]
        g, S, Pp, A, f = dall(a, kmersize=15, skipcost=40., maxdiff=50, maxgap=1000)
        out["g_%d_a" % ci] = a.astype(np.int32) if a.max() < 2**31 else a
        out["g_%d_exact" % ci] = np.array(g)
        if g >= 0:
            out["g_%d_S" % ci] = S
            out["g_%d_P" % ci] = Pp
            out["g_%d_A" % ci] = A
        g, S, Pp, A = dfast(a, kmersize=15, skipcost=40., maxdiff=50, maxgap=1000)
        out["g_%d_fg" % ci] = np.array(g)
        out["g_%d_fS" % ci] = S
        out["g_%d_fP" % ci] = Pp
        out["g_%d_fA" % ci] = A
    out["g_count"] = np.array(len(cases))
    # --- local DP cases ---
    fl = getattr(m, P)
    flm = getattr(m, P + "_mismatch")
    flf = getattr(m, P + "_fast")
    flmf = getattr(m, P + "_mismatch_fast")
    lc = []
    for t in range(10):
        if t % 4 == 0:
            a = synth.anchors_tieheavy(rng, n=int(rng.integers(3, 400)), k=9)
        else:
            a = synth.anchors_local(rng, n_true=int(rng.integers(5, 900)), n_noise=int(rng.integers(0, 300)), multi=(t % 2 == 0))
        lc.append(a)
    for ci, a in enumerate(lc):
        a = a[nb_argsort(a[:, 0] + a[:, 3])]
        out["l_%d_a" % ci] = a.astype(np.int32)
        for tag, f, sk, mg in (("fl", fl, 40., 99), ("flm", flm, 40., 99), ("fl59", fl, 59., 50),
                               ("flf", flf, 40., 99), ("flmf", flmf, 40., 99)):
            sc, path = f(a, kmersize=9, skipcost=sk, maxdiff=30, maxgap=mg)
            out["l_%d_%s_score" % (ci, tag)] = np.array(sc)
            out["l_%d_%s_path" % (ci, tag)] = np.array(path, dtype=np.int64).astype(np.int32)
    out["l_count"] = np.array(len(lc))
    np.savez_compressed(os.path.join(HERE, "chain.npz"), **out)
    print("chain.npz:", os.path.getsize(os.path.join(HERE, "chain.npz")), "bytes")


def gen_asm_linked():
    """asm mode: the reference's linked global DP with carry-in, run batch after batch inside the
    first-round loop of `assembly_get_readmap_DP_test` (oracle/asm.py transcribes that inline loop; the DP called
    here is the reference's own njit function) -> tests/golden/asm_linked.npz."""
    import numba
    import oracle.asm as oasm
    m = refimport.load_mode("asm")
    dall = getattr(m, "linked_" + P + "_d_all")
    dfast = getattr(m, "linked_" + P + "_d_fast_all")

    @numba.njit
    def nb_argsort(a):
        return np.argsort(a)

    rng = np.random.default_rng(20261018)
    out = {}
    flows = []
    for t in range(7):
        if t == 0:
            a = synth.anchors_tieheavy(rng, n=600)
        elif t == 1:
            a = synth.anchors_global(rng, n_true=40, n_noise=0)
        else:
            a = synth.anchors_global(rng, n_true=int(rng.integers(200, 900)), n_noise=int(rng.integers(0, 1200)))
        a = a[nb_argsort(a[:, 0])].astype(np.int64)
        nb = int(rng.integers(2, 6)) if t != 1 else 1
        cuts = sorted(set(int(c) for c in rng.integers(1, len(a) - 1, size=nb - 1))) if nb > 1 else []
        # a batch boundary never splits a run of equal read positions (batches are read-position slices, asm:22415-22441)
        def slide(c):
            while c < len(a) and a[c][0] == a[c - 1][0]:
                c += 1
            return c
        cuts = sorted(set(c for c in map(slide, cuts) if c < len(a)))
        flows.append(np.split(a, cuts))
    # an empty batch in the middle, and a flow whose first batch is a single anchor (P[g] < 0: nothing carried)
    flows.append([flows[2][0], np.zeros((0, 4), np.int64)] + flows[2][1:])
    flows.append([flows[3][0][:1]] + [flows[3][0][1:]] + flows[3][1:])
    # a noise-only batch after a normal one: the exact DP bails out on opcount (:21754) and the loop falls back to
    # the heuristic twin with the carried prefix (:23246-23247); then a normal batch again
    noise = synth.anchors_global(rng, n_true=0, n_noise=3200, repeats=False)
    noise = noise[nb_argsort(noise[:, 0])].astype(np.int64)
    shift = int(flows[5][0][:, 0].max()) + 1
    noise[:, 0] += shift
    tail = flows[5][1].copy()
    tail[:, 0] += int(noise[:, 0].max()) + 1
    flows.append([flows[5][0], noise, tail])
    for fi, batches in enumerate(flows):
        rec = []

        def dp(gs, gi, pS, pP, prl, lk, rec=rec):
            g, S, Pp, A, _ = dall(gs, gi, pS, pP, prl, lk, kmersize=15, skipcost=40., maxdiff=50, maxgap=1000)
            rec.append((np.array([gs, gi, prl], dtype=np.float64), pS.copy(), pP.copy(), lk.copy(), int(g), S.copy(), Pp.copy(),
                        A.copy()))
            return g, S, Pp, A

        def dpf(gs, gi, pS, pP, prl, lk, rec=rec):
            g, S, Pp, A = dfast(gs, gi, pS, pP, prl, lk, kmersize=15, skipcost=40., maxdiff=50, maxgap=1000)
            rec[-1] = rec[-1] + (int(g), S.copy(), Pp.copy(), A.copy())
            return g, S, Pp, A

        path = oasm.first_round_path(batches, 15, 40., 50, 1000, dp=dp, dp_fast=dpf)
        out["f%d_nb" % fi] = np.array(len(batches))
        for bi, b in enumerate(batches):
            out["f%d_b%d" % (fi, bi)] = b.astype(np.int32) if (len(b) == 0 or b.max() < 2**31) else b
        out["f%d_calls" % fi] = np.array(len(rec))
        for ci, r in enumerate(rec):
            hd, pS, pP, lk, g, S, Pp, A = r[:8]
            # the heuristic twin on the same arguments: what the loop used after a bail-out, else run here
            fg, fS, fP, fA = r[8:] if len(r) > 8 else dfast(hd[0] if len(pS) else 0, int(hd[1]) if len(pS) else 0, pS, pP, int(hd[2]),
                                                            lk, kmersize=15, skipcost=40., maxdiff=50, maxgap=1000)
            out["f%d_c%d_fg" % (fi, ci)] = np.array(int(fg))
            out["f%d_c%d_fS" % (fi, ci)] = fS
            out["f%d_c%d_fP" % (fi, ci)] = fP
            out["f%d_c%d_fA" % (fi, ci)] = fA
            out["f%d_c%d_head" % (fi, ci)] = hd
            out["f%d_c%d_preS" % (fi, ci)] = pS
            out["f%d_c%d_preP" % (fi, ci)] = pP
            out["f%d_c%d_g" % (fi, ci)] = np.array(g)
            out["f%d_c%d_S" % (fi, ci)] = S
            out["f%d_c%d_P" % (fi, ci)] = Pp
            out["f%d_c%d_A" % (fi, ci)] = A
        out["f%d_path" % fi] = np.array(path, dtype=np.int64).reshape(-1, 4)
        print("flow", fi, "batches", [len(b) for b in batches], "calls", len(rec), "path", len(path))
    out["n_flows"] = np.array(len(flows))
    # --- second round: linked_..._fine_list_all over local (k = 9) anchors, sorted by read start ---
    dloc = getattr(m, "linked_" + P + "_all")
    lflows = []
    for t in range(6):
        if t == 0:
            a = synth.anchors_tieheavy(rng, n=500, k=9)
        else:
            a = synth.anchors_local(rng, n_true=int(rng.integers(100, 900)), n_noise=int(rng.integers(0, 300)), multi=(t % 2 == 0))
        a = a[nb_argsort(a[:, 0])].astype(np.int64)
        nb = int(rng.integers(1, 5))
        cuts = sorted(set(int(c) for c in rng.integers(1, len(a) - 1, size=nb - 1))) if nb > 1 else []

        def slide(c):
            while c < len(a) and a[c][0] == a[c - 1][0]:
                c += 1
            return c
        cuts = sorted(set(c for c in map(slide, cuts) if c < len(a)))
        lflows.append(np.split(a, cuts))
    for fi, batches in enumerate(lflows):
        rec = []

        def dpl(gs, gi, pS, pP, prl, lk, rec=rec):
            g, S, Pp, A, _ = dloc(gs, gi, pS, pP, prl, lk, kmersize=9, skipcost=30., maxdiff=30, maxgap=99)
            rec.append((int(g), S.copy(), Pp.copy(), A.copy()))
            return g, S, Pp, A

        try:
            path = oasm.second_round_path(batches, 9, 30., 30, 99, dp=dpl)
            out["l%d_err" % fi] = np.array(0)
        except IndexError:
            # a chain that STARTS in a carried anchor: pre_P = -(-9999999) is followed as an index by the traceback
            # (:23380-23385) -- the reference raises here too and the contig yields nothing
            path = []
            out["l%d_err" % fi] = np.array(1)
        out["l%d_nb" % fi] = np.array(len(batches))
        for bi, b in enumerate(batches):
            out["l%d_b%d" % (fi, bi)] = b.astype(np.int32) if (len(b) == 0 or b.max() < 2**31) else b
        out["l%d_calls" % fi] = np.array(len(rec))
        for ci, (g, S, Pp, A) in enumerate(rec):
            out["l%d_c%d_g" % (fi, ci)] = np.array(g)
            out["l%d_c%d_S" % (fi, ci)] = S
            out["l%d_c%d_P" % (fi, ci)] = Pp
            out["l%d_c%d_A" % (fi, ci)] = A
        out["l%d_path" % fi] = np.array(path, dtype=np.int64).reshape(-1, 4)
        print("local flow", fi, "batches", [len(b) for b in batches], "calls", len(rec), "path", len(path))
    out["n_lflows"] = np.array(len(lflows))
    np.savez_compressed(os.path.join(HERE, "asm_linked.npz"), **out)
    print("asm_linked.npz:", os.path.getsize(os.path.join(HERE, "asm_linked.npz")), "bytes")


def asm_reseed_inputs():
    """Seeded inputs of the asm re-seeding fixture (shared with tests/test_oracle_asm.py): a 2-contig reference and
    one 60 kb 'contig read' with an inversion and a deletion, 1 % divergence."""
    ref = synth.make_reference(77, 240000, n_contigs=2)
    rng = np.random.default_rng(78)
    src = np.frombuffer(ref[0][1].encode(), dtype=np.uint8)[20000:82000].copy()
    comp = np.zeros(256, np.uint8)
    for x, y in zip(b"ACGT", b"TGCA"):
        comp[x] = y
    parts = [src[:20000], comp[src[20000:26000]][::-1], src[26000:40000], src[43000:]]      # INV 6 kb, DEL 3 kb
    read = synth.mutate(rng, np.concatenate(parts), 0.01)
    return ref, read.tobytes().decode()


def gen_asm_reseed():
    """asm mode: the reference's yield_second_mapinfo / collect_second_round_anchors (mammap_asm.py:22444-22755) on a
    first-round path computed by the oracle -> tests/golden/asm_reseed.npz."""
    import oracle
    import oracle.asm as oasm
    import oracle.pipeline as pl
    import refrun
    ref, read = asm_reseed_inputs()
    R = refrun.ReferenceRunner(ref, mode="asm")
    ox = oracle.Index(ref)
    a = np.array(ox.map(read, -1, -1), dtype=np.int64)
    a = a[oracle.argsort_i64(a[:, 0])]
    path = oasm.first_round_path([a], 15, 40., 50, 1000)
    raw = np.array(path[::-1], dtype=np.int64)
    from vacmap_b200.sam import reverse_complement
    rc = reverse_complement(read)
    out = {"raw": raw}
    for bi, batch in enumerate((8000, 20000, 100000)):
        got = [np.asarray(x) for x in R.mod.yield_second_mapinfo(raw, read, rc, R.contig2start, R.contig2seq, 9, batch)]
        out["n_%d" % bi] = np.array(len(got))
        out["batch_%d" % bi] = np.array(batch)
        for ci, x in enumerate(got):
            out["b%d_%d" % (bi, ci)] = x.astype(np.int64)
        print("batch", batch, "->", [len(x) for x in got])
    np.savez_compressed(os.path.join(HERE, "asm_reseed.npz"), **out)
    print("asm_reseed.npz:", os.path.getsize(os.path.join(HERE, "asm_reseed.npz")), "bytes")


def asm_e2e_inputs():
    return synth.asm_e2e_inputs()


def gen_asm_e2e():
    """asm mode end to end: the reference's assembly_get_readmap_DP_test (mammap_asm.py:23204) over the oracle
    natives on a 520 kb contig read -> tests/golden/asm_e2e.json.gz (onemapinfolist rows)."""
    import gzip
    import json
    import tempfile
    import time
    import refrun
    from vacmap_b200.sam import reverse_complement
    ref, read = asm_e2e_inputs()
    assert len(read) >= 500000
    out = {"cases": []}
    for eqx in (False, True):
        R = refrun.ReferenceRunner(ref, mode="asm", eqx=eqx)
        wd = tempfile.mkdtemp() + "/w/"
        t0 = time.time()
        recs = R.mod.assembly_get_readmap_DP_test(wd, "ctgread", read, reverse_complement(read), len(read), R.aligner,
                                                  R.mod.pos2contig, R.contig2start, R.contig2seq, R.index2contig, R.option)
        recs = [list(r) for r in recs]
        for r in recs:
            for i in (3, 4, 5, 6, 7):
                r[i] = int(r[i])
        print("eqx", eqx, len(recs), "records", [(r[2], r[3], r[4], r[5], r[6]) for r in recs], round(time.time() - t0, 1), "s")
        out["cases"].append({"eqx": eqx, "records": recs})
    with gzip.open(os.path.join(HERE, "asm_e2e.json.gz"), "wt") as f:
        json.dump(out, f)
    print("asm_e2e.json.gz:", os.path.getsize(os.path.join(HERE, "asm_e2e.json.gz")), "bytes")


def gen_asm_e2e_2():
    """A second contig through the reference's assembly_get_readmap_DP_test: reverse strand, a translocated piece of the
    other contig, a tandem duplication and a deletion (synth.asm_e2e_inputs_2) -> tests/golden/asm_e2e2.json.gz."""
    import gzip
    import json
    import tempfile
    import time
    import refrun
    from vacmap_b200.sam import reverse_complement
    ref, read = synth.asm_e2e_inputs_2()
    assert len(read) >= 500000
    out = {"cases": []}
    for eqx in (True,):
        R = refrun.ReferenceRunner(ref, mode="asm", eqx=eqx)
        wd = tempfile.mkdtemp() + "/w/"
        t0 = time.time()
        recs = R.mod.assembly_get_readmap_DP_test(wd, "ctgread2", read, reverse_complement(read), len(read), R.aligner,
                                                  R.mod.pos2contig, R.contig2start, R.contig2seq, R.index2contig, R.option)
        recs = [list(r) for r in recs]
        for r in recs:
            for i in (3, 4, 5, 6, 7):
                r[i] = int(r[i])
        print("eqx", eqx, len(recs), "records", [(r[1], r[2], r[3], r[4], r[5], r[6]) for r in recs], round(time.time() - t0, 1), "s")
        out["cases"].append({"eqx": eqx, "records": recs})
    with gzip.open(os.path.join(HERE, "asm_e2e2.json.gz"), "wt") as f:
        json.dump(out, f)
    print("asm_e2e2.json.gz:", os.path.getsize(os.path.join(HERE, "asm_e2e2.json.gz")), "bytes")


def gen_asm_link():
    """link_cigar (mammap_asm.py:22366-22410, njit) on random CIGAR pairs -> tests/golden/asm_link_cigar.json."""
    import json
    m = refimport.load_mode("asm")
    rng = np.random.default_rng(5)

    def rnd():
        n = int(rng.integers(1, 5))
        ops, last = [], ""
        for _ in range(n):
            op = str(rng.choice([c for c in "=XIDM" if c != last]))
            last = op
            ops.append(str(int(rng.choice([1, 7, 10, 99, 100, 1234, 20000]))) + op)
        return "".join(ops)
    rows = []
    for _ in range(300):
        a, b = rnd(), rnd()
        rows.append([a, b, str(m.link_cigar(a, b))])
    with open(os.path.join(HERE, "asm_link_cigar.json"), "w") as f:
        json.dump(rows, f)
    print("asm_link_cigar.json", len(rows), "pairs,", sum(1 for r in rows if r[2] != r[0] + r[1]), "merged")


def squash_sam_line(line):
    """SEQ / QUAL of a 520 kb contig read do not belong in a fixture: replaced by length + sha1."""
    import hashlib
    f = line.split("\t")
    for i in (9, 10):
        if len(f[i]) > 64:
            f[i] = "%d:%s" % (len(f[i]), hashlib.sha1(f[i].encode()).hexdigest())
    return "\t".join(f)


def gen_asm_sam():
    """asm mode's SAM emitter iterator_get_bam_dict_str (mammap_asm.py:22757-22941) on the records of asm_e2e.json.gz
    (plus a MAPQ 1 / 0 variant for the primary rule) -> tests/golden/asm_sam.json.gz."""
    import gzip
    import json
    import refrun
    E = json.load(gzip.open(os.path.join(HERE, "asm_e2e.json.gz"), "rt"))
    ref, read = asm_e2e_inputs()
    R = refrun.ReferenceRunner(ref, mode="asm")
    qual = "".join(chr(33 + (i * 11) % 41) for i in range(len(read)))
    out = []
    variants = [
        dict(eqx=False, md=False, H=False, fakecigar=False, mapq=None, qual=False),
        dict(eqx=True, md=True, H=False, fakecigar=False, mapq=None, qual=True),
        dict(eqx=True, md=True, H=True, fakecigar=True, mapq=None, qual=True, shortcs=False),
        dict(eqx=False, md=False, H=True, fakecigar=False, mapq=[1, 60, 0], qual=False),
    ]
    for v in variants:
        recs = [list(r) for r in [c for c in E["cases"] if c["eqx"] == v["eqx"]][0]["records"]]
        if v["mapq"]:
            # longest record MAPQ 1 -> the second longest becomes primary; a MAPQ 0 record is written as 1
            order = sorted(range(len(recs)), key=lambda i: recs[i][4] - recs[i][3])[::-1]
            for rank, i in enumerate(order):
                recs[i][7] = v["mapq"][rank]
        if v["H"]:
            # the rows were assembled with soft clips: re-clip for the hard-clip variant
            for r in recs:
                r[8] = r[8].replace("S", "H")
        opt = dict(R.option)
        opt.update({"H": v["H"], "fakecigar": v["fakecigar"]})
        lines = list(R.mod.iterator_get_bam_dict_str([tuple(r) for r in recs], read.upper(), qual if v["qual"] else None,
                                                     R.contig2iloc, R.contig2seq, v["md"], v.get("shortcs", True), False, False, opt))
        out.append({"variant": v, "records": recs, "sam": [squash_sam_line(x) for x in lines]})
        print(v, len(lines), "lines", [x.split("\t")[1] + ":" + x.split("\t")[4] for x in lines])
    with gzip.open(os.path.join(HERE, "asm_sam.json.gz"), "wt") as f:
        json.dump(out, f)
    print("asm_sam.json.gz:", os.path.getsize(os.path.join(HERE, "asm_sam.json.gz")), "bytes")


def gen_e2e():
    """End-to-end records from the reference's own get_readmap_DP_test / get_bam_dict_str run over the
    oracle's vacmap_index / edlib shim.  Inputs are regenerated from seeds (tests/synth.py) except the
    reference's testdata pair, which is copied as a fixture (BASELINE configs[0])."""
    import gzip
    import json
    import shutil
    import refrun
    import oracle.shim as shim
    td = os.path.join(HERE, "testdata")
    os.makedirs(td, exist_ok=True)
    for f in ("read.fasta", "reference.fasta"):
        with open(os.path.join(refimport.REF_ROOT, "testdata", f), "rb") as fi, gzip.open(os.path.join(td, f + ".gz"), "wb") as fo:
            shutil.copyfileobj(fi, fo)
    out = {"cases": []}

    def case(name, ref, reads, mode, **opt):
        R = refrun.ReferenceRunner(ref, mode=mode, **opt)
        recs, sams = [], []
        for rid, seq in reads:
            r = R.align(rid, seq)
            recs.append([list(x) for x in r])
            sams.append(R.mod.get_bam_dict_str(r, seq.upper(), None, R.contig2iloc, R.contig2seq, R.option["md"],
                                               R.option["shortcs"], R.option["cigar2cg"], R.option["markunbalancetra"],
                                               R.option) if r else [])
        out["cases"].append({"name": name, "mode": mode, "opt": opt, "records": recs, "sam": sams})
        print(name, "reads", len(reads), "records", sum(len(r) for r in recs))

    ref = [(n, s) for n, s, _ in shim.read_fastx(os.path.join(td, "reference.fasta.gz"))]
    reads = [(r[0], r[1]) for r in shim.read_fastx(os.path.join(td, "read.fasta.gz"))]
    case("testdata_H", ref, reads, "H")
    case("testdata_H_eqx_md", ref, reads, "H", eqx=True, md=True)
    ref2 = synth.make_reference(1, 300000)
    reads2 = synth.make_reads(ref2, 11, 16, read_len=6000, err=0.10, sv_frac=0.5)
    case("synth300k_H", ref2, reads2, "H")
    case("synth300k_H_eqx", ref2, reads2[:6], "H", eqx=True, md=True)
    ref3 = synth.make_reference(3, 600000, n_contigs=2)
    reads3 = synth.make_reads(ref3, 12, 8, read_len=15000, err=0.10, sv_frac=0.3)
    case("synth600k_2ctg_H", ref3, reads3, "H")
    # more of the option / mode space, all from the reference's own code: hard clips with the approximate SA CIGAR,
    # mode L (mammap_ccs) on HiFi-like reads, mode S (mammap_sensitive) on reads with nested SVs
    case("synth300k_H_hardclip_fakecigar", ref2, reads2[:8], "H", H=True, fakecigar=True)
    reads4 = synth.make_reads(ref2, 13, 8, read_len=6000, err=0.005, ratio=(1, 1, 1), sv_frac=0.5)
    case("synth300k_L", ref2, reads4, "L")
    case("synth300k_L_eqx_md_longcs", ref2, reads4[:4], "L", eqx=True, md=True, shortcs=False)
    reads5 = synth.make_reads(ref2, 14, 8, read_len=6000, err=0.10, sv_frac=0.8)
    case("synth300k_S", ref2, reads5, "S")
    with gzip.open(os.path.join(HERE, "e2e.json.gz"), "wt") as f:
        json.dump(out, f)
    print("e2e.json.gz:", os.path.getsize(os.path.join(HERE, "e2e.json.gz")), "bytes")


if __name__ == "__main__":
    what = sys.argv[1:] or ["chain"]
    if "chain" in what:
        gen_chain()
    if "e2e" in what:
        gen_e2e()
    if "asm" in what:
        gen_asm_linked()
    if "asmseed" in what:
        gen_asm_reseed()
    if "asme2e" in what:
        gen_asm_e2e()
    if "asme2e2" in what:
        gen_asm_e2e_2()
    if "asmlink" in what:
        gen_asm_link()
    if "asmsam" in what:
        gen_asm_sam()


def gen_sam_comments():
    """SAM text with FASTQ comments and base qualities (get_bam_dict_str_comments, --copycomments) from the reference's
    own emitter, fed with the records already pinned in e2e.json.gz -> tests/golden/sam_comments.json.gz."""
    import gzip
    import json
    import refrun
    sys.path.insert(0, os.path.dirname(HERE))
    from test_oracle_e2e import case_inputs
    e2e = json.load(gzip.open(os.path.join(HERE, "e2e.json.gz"), "rt"))
    out = {"cases": []}
    comments = "XA:i:5\tNM:i:7\tab:Z:hello there\tbad\tzz:Q:1\tRG:Z:other\tXB:f:1.5"
    for name in ("testdata_H_eqx_md", "synth300k_H", "synth300k_H_hardclip_fakecigar"):
        case = [c for c in e2e["cases"] if c["name"] == name][0]
        ref, reads = case_inputs(name)
        R = refrun.ReferenceRunner(ref, mode=case["mode"], **case["opt"])
        lines = []
        for (rid, seq), recs in zip(reads, case["records"]):
            if not recs:
                lines.append([])
                continue
            qual = "".join(chr(33 + (i * 7) % 40) for i in range(len(seq)))
            lines.append(R.mod.get_bam_dict_str_comments([tuple(r) for r in recs], seq.upper(), qual, comments, R.contig2iloc,
                                                         R.contig2seq, R.option["md"], R.option["shortcs"], R.option["cigar2cg"],
                                                         R.option["markunbalancetra"], R.option))
        out["cases"].append({"name": name, "comments": comments, "sam": lines})
        print(name, sum(len(x) for x in lines), "lines")
    with gzip.open(os.path.join(HERE, "sam_comments.json.gz"), "wt") as f:
        json.dump(out, f)


if __name__ == "__main__" and "samcomments" in sys.argv[1:]:
    gen_sam_comments()
