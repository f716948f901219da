"""Oracle per-read pipeline (C natives + Python glue) vs records produced by the REFERENCE's own
get_readmap_DP_test run over the same natives (tests/golden/e2e.json.gz)."""
import gzip
import json
import os

import pytest

import oracle
import oracle.pipeline as pl
import oracle.shim as shim
import synth

HERE = os.path.dirname(os.path.abspath(__file__))
E2E = json.load(gzip.open(os.path.join(HERE, "golden", "e2e.json.gz"), "rt"))


def case_inputs(name):
    td = os.path.join(HERE, "golden", "testdata")
    if name.startswith("testdata"):
        ref = [(n, s) for n, s, _ in shim.read_fastx(os.path.join(td, "reference.fasta.gz"))]
        reads = [(r[0], r[1]) for r in shim.read_fastx(os.path.join(td, "read.fasta.gz"))]
    elif name.startswith("synth300k_L"):
        ref = synth.make_reference(1, 300000)
        reads = synth.make_reads(ref, 13, 8, read_len=6000, err=0.005, ratio=(1, 1, 1), sv_frac=0.5)
        if name.endswith("longcs"):
            reads = reads[:4]
    elif name == "synth300k_S":
        ref = synth.make_reference(1, 300000)
        reads = synth.make_reads(ref, 14, 8, read_len=6000, err=0.10, sv_frac=0.8)
    elif name.startswith("synth300k"):
        ref = synth.make_reference(1, 300000)
        reads = synth.make_reads(ref, 11, 16, read_len=6000, err=0.10, sv_frac=0.5)
        if name.endswith("eqx"):
            reads = reads[:6]
        if name.endswith("fakecigar"):
            reads = reads[:8]
    elif name.startswith("synth600k"):
        ref = synth.make_reference(3, 600000, n_contigs=2)
        reads = synth.make_reads(ref, 12, 8, read_len=15000, err=0.10, sv_frac=0.3)
    else:
        raise KeyError(name)
    return ref, reads


def option_for(case):
    import refrun_options
    return refrun_options.default_option(case["mode"], **case["opt"])


@pytest.mark.parametrize("ci", range(len(E2E["cases"])))
def test_oracle_pipeline_matches_reference_records(ci):
    case = E2E["cases"][ci]
    ref, reads = case_inputs(case["name"])
    opt = option_for(case)
    ix = oracle.Index(ref, w=10, k=15)
    ctg = pl.Contigs([n for n, _ in ref], [s for _, s in ref])
    assert len(reads) == len(case["records"])
    for (rid, seq), want in zip(reads, case["records"]):
        got = pl.align_read(rid, seq, ix, ctg, opt, case["mode"])
        assert [list(r) for r in got] == want, rid


def test_testdata_gives_three_alignments():
    """README.md:124 of the reference: the testdata pair yields three alignments (+, -, +)."""
    case = E2E["cases"][0]
    recs = case["records"][0]
    assert len(recs) == 3 and [r[2] for r in recs] == ["+", "-", "+"]
