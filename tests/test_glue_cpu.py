"""Product host driver + glue (vm_pipeline.hpp / vm_glue.hpp) checked on CPU: the harness in
tests/gluetest plugs the ORACLE's C stage functions in place of the CUDA kernels, and the records
must equal the ones the reference's own Python produced (tests/golden/e2e.json.gz)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
from test_oracle_e2e import E2E, case_inputs, option_for

HERE = os.path.dirname(os.path.abspath(__file__))
OPS = "MIDNSHP=X"


class GtOptions(ctypes.Structure):
    _fields_ = [("global_skipcost", ctypes.c_double), ("local_skipcost", ctypes.c_double),
                ("maxdivergence", ctypes.c_double), ("accept", ctypes.c_double)] + \
               [(n, ctypes.c_int32) for n in ("global_maxdiff", "local_maxdiff", "check_num", "eqx", "hardclip", "nodiscard",
                                              "max_guides", "local_maxgap", "clamp40", "kmersize", "threads")]


MODE = {"H": (60.0, 5, 99, 0), "L": (40.0, 3, 50, 1), "S": (40.0, 0, 99, 0)}


def build_harness():
    oracle.build()
    d = os.path.join(HERE, "gluetest")
    out = os.path.join(d, "libgluetest.so")
    srcs = [os.path.join(d, "gluetest.cpp")] + [os.path.join(HERE, "..", "vacmap_b200", "csrc", f) for f in ("vm_glue.hpp", "vm_pipeline.hpp", "vm_dglue.hpp", "vm_dgrun.hpp")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, srcs[0],
                               "-L" + os.path.join(HERE, "..", "oracle", "_build"), "-loracle",
                               "-Wl,-rpath," + os.path.abspath(os.path.join(HERE, "..", "oracle", "_build")), "-lpthread"])
    L = ctypes.CDLL(out)
    L.gt_align_batch.restype = ctypes.c_int64
    return L


def glue_align(L, ref, reads, opt, mode, threads=4, k=15, w=10, device_glue=False, counters=None):
    """device_glue: extend_func through the device-resident glue (vm_dgrun.hpp / vm_dglue.hpp, the code the CUDA kernels
    run) instead of the vector-based host glue; counters: dict that receives the branch counters and per-read status."""
    import tempfile
    os.environ.pop("GT_DEVICE_GLUE", None)
    os.environ.pop("GT_COUNTERS", None)
    if device_glue:
        os.environ["GT_DEVICE_GLUE"] = "1"
    cfile = None
    if counters is not None:
        cfile = tempfile.mktemp(suffix=".gt")
        os.environ["GT_COUNTERS"] = cfile
    ix = oracle.Index(ref, w=w, k=k)
    t = oracle.tables()
    cat = "".join(s.upper() for _, s in ref).encode()
    starts = np.array(ix.offsets[:-1], dtype=np.int64)
    lens = np.diff(ix.offsets).astype(np.int64)
    rcat = "".join(s.upper() for _, s in reads).encode()
    roff = np.zeros(len(reads) + 1, np.int64)
    for i, (_, s) in enumerate(reads):
        roff[i + 1] = roff[i] + len(s)
    acc, mg, lmg, c40 = MODE[mode]
    o = GtOptions(opt["golbal_skipcost"], opt["local_skipcost"], opt["maxdivergence"], acc, opt["golbal_maxdiff"],
                  opt["local_maxdiff"], opt["c"], int(opt["eqx"]), int(opt["H"]), int(opt["nodiscard"]), mg, lmg, c40, k, threads)
    rec_cap, cig_cap = 64 * len(reads) + 64, 4_000_000
    rows = np.zeros((rec_cap, 9), np.int64)
    cig = np.zeros(cig_cap, np.uint32)
    ncig = ctypes.c_int64(0)
    vp = ctypes.c_void_p
    n = L.gt_align_batch(vp(ix.h), vp(ctypes.addressof(t["struct"])), cat, vp(starts.ctypes.data), vp(lens.ctypes.data),
                         ctypes.c_int32(len(ref)), rcat, vp(roff.ctypes.data), ctypes.c_int64(len(reads)), ctypes.byref(o),
                         vp(rows.ctypes.data), ctypes.c_int64(rec_cap), vp(cig.ctypes.data), ctypes.c_int64(cig_cap),
                         ctypes.byref(ncig))
    assert n <= rec_cap and ncig.value <= cig_cap
    os.environ.pop("GT_DEVICE_GLUE", None)
    os.environ.pop("GT_COUNTERS", None)
    if cfile is not None:
        counters["status"] = {}
        for line in open(cfile):
            f = line.split()
            if f[0] == "status":
                counters["status"][int(f[1])] = int(f[2])
            else:
                counters[f[0]] = int(f[1])
        os.unlink(cfile)
    out = [[] for _ in reads]
    co = 0
    for r in rows[:n]:
        ops = cig[co:co + r[8]]
        co += int(r[8])
        s = "".join("%d%s" % (int(x) >> 4, OPS[int(x) & 0xf]) for x in ops)
        rid = int(r[0])
        out[rid].append([reads[rid][0], ref[int(r[1])][0], "+" if r[2] == 1 else "-", int(r[3]), int(r[4]), int(r[5]),
                         int(r[6]), int(r[7]), s])
    return out


@pytest.fixture(scope="module")
def harness():
    return build_harness()


@pytest.mark.parametrize("ci", range(len(E2E["cases"])))
def test_glue_matches_reference_records(harness, ci):
    case = E2E["cases"][ci]
    ref, reads = case_inputs(case["name"])
    got = glue_align(harness, ref, reads, option_for(case), case["mode"])
    for (rid, _), g, w in zip(reads, got, case["records"]):
        assert g == w, rid


@pytest.mark.parametrize("ci", range(len(E2E["cases"])))
def test_device_glue_matches_reference_records(harness, ci):
    """The same per-read functions and launch sequence the CUDA kernels run (vm_dglue.hpp / vm_dgrun.hpp), in host loops
    over the oracle natives: records equal to the reference's."""
    case = E2E["cases"][ci]
    ref, reads = case_inputs(case["name"])
    got = glue_align(harness, ref, reads, option_for(case), case["mode"], device_glue=True)
    for (rid, _), g, w in zip(reads, got, case["records"]):
        assert g == w, rid


@pytest.mark.parametrize("device_glue", [False, True])
@pytest.mark.parametrize("name", ["bulk_H", "bulk_H_eqx", "bulk_L_k19", "bulk_S"])
def test_glue_on_bulk_sample(harness, name, device_glue):
    """Both glue implementations on a sample of the bulk fixture (reference-generated, tests/golden/bulk_e2e.json.gz)
    that contains every counted branch: records, branch counters and the per-read status."""
    import bulk
    import refrun_options
    from test_oracle_bulk import BULK, BRANCHES, sample_of
    case = BULK["cases"][name]
    ref = bulk.reference_for(name)
    allreads = bulk.reads_for(name, ref)
    pick = sample_of(case)
    reads = [allreads[i] for i in pick]
    opt = refrun_options.default_option(case["mode"], **case["opt"])
    cnt = {}
    got = glue_align(harness, ref, reads, opt, case["mode"], k=case["k"], w=case["w"], device_glue=device_glue, counters=cnt)
    want_tot = {b: 0 for b in BRANCHES}
    for j, i in enumerate(pick):
        want = case["reads"][i]
        assert [bulk.squash(r) for r in got[j]] == want["records"], want["id"]
        for b in BRANCHES:
            want_tot[b] += want["cnt"].get(b, 0)
        st = cnt["status"][j]
        assert (st == 0) == (want["status"] == "ok"), (want["id"], st)
        if want["status"] == "unmapped":
            assert st in (1, 2, 3, 5), (want["id"], st)
    for b in BRANCHES:
        assert cnt.get("c_" + b, 0) == want_tot[b], b
