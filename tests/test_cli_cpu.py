"""The command line's host plumbing on a box without a GPU: `vacmap_b200.__main__.main` with the index and the aligner
replaced by stand-ins that hand back the REFERENCE's records (tests/golden/e2e.json.gz).  Everything else is the product's:
the reader, batching, the read-name filter, the native SAM emitter (vm_sam_batch), header, SAM / BAM output."""
import gzip
import os
import shutil

import numpy as np
import pytest

from test_oracle_e2e import E2E, case_inputs, option_for
from test_sam_native import pack_records
import vacmap_b200.__main__ as cli
from vacmap_b200 import align, bam

HERE = os.path.dirname(os.path.abspath(__file__))


class FakeIndex:
    h = None

    def __init__(self, ref, w=10, k=15, device=0, **_kw):
        ref = FakeIndex.contigs
        self.names = [n for n, _ in ref]
        self._seqs = dict(ref)
        self.lens = [len(s) for _, s in ref]
        self.k, self.w = k, w

    def seq(self, name, start=0, end=0x7fffffff):
        return self._seqs[name][start:end]

    def write_mmi(self, path):
        pass

    def close(self):
        pass


class FakeAligner:
    """submit_packed / wait with the batch's golden rows, looked up by read sequence."""
    rows_by_seq = {}

    def __init__(self, index, opt, mode, host_threads=0, **_kw):
        self.index = index
        self.last_stage_ms = {"total": 1.0, "k_fill": 0.5, "n_fill_jobs": 3.0}

    def submit_packed(self, seq_cat, seq_off, resident=False):
        seqs = [bytes(seq_cat[seq_off[i]:seq_off[i + 1]]).decode() for i in range(len(seq_off) - 1)]
        return seqs

    def wait(self, handle):
        return pack_records([FakeAligner.rows_by_seq.get(s, []) for s in handle], self.index.names)


@pytest.fixture
def fake_gpu(monkeypatch):
    monkeypatch.setattr(align, "Index", FakeIndex)
    monkeypatch.setattr(align, "Aligner", FakeAligner)


def _write_inputs(tmp_path, ref, reads, fastq=False):
    (tmp_path / "ref.fa").write_text("".join(">%s\n%s\n" % (n, s) for n, s in ref))
    if fastq:
        (tmp_path / "reads.fq").write_text("".join("@%s\n%s\n+\n%s\n" % (n, s, "I" * len(s)) for n, s in reads))
        return str(tmp_path / "reads.fq")
    (tmp_path / "reads.fa").write_text("".join(">%s\n%s\n" % (n, s) for n, s in reads))
    return str(tmp_path / "reads.fa")


@pytest.mark.parametrize("ci", range(len(E2E["cases"])))
def test_command_line_writes_the_references_sam_lines(tmp_path, fake_gpu, ci):
    case = E2E["cases"][ci]
    ref, reads = case_inputs(case["name"])
    opt = option_for(case)
    FakeIndex.contigs = ref
    FakeAligner.rows_by_seq = {seq.upper(): [tuple(r) for r in recs] for (_, seq), recs in zip(reads, case["records"])}
    rpath = _write_inputs(tmp_path, ref, reads)
    flags = []
    if opt.get("eqx"):
        flags.append("--eqx")
    if opt.get("md"):
        flags.append("--MD")
        flags += ["--cs"] if opt.get("shortcs", True) else ["--cs=long"]
    if opt.get("H"):
        flags.append("--H")
    if opt.get("fakecigar"):
        flags.append("--fakecigar")
    out = tmp_path / "out.sam"
    cli.main(["-ref", str(tmp_path / "ref.fa"), "-read", rpath, "-mode", case["mode"], "--nowriteindex", "--batch-bases", "20000",
              "-o", str(out)] + flags)
    lines = out.read_text().splitlines()
    body = [l for l in lines if not l.startswith("@")]
    assert lines[0] == "@HD\tVN:1.0" and sum(l.startswith("@SQ") for l in lines) == len(ref)
    assert any(l.startswith("@RG\tID:1\tSM:sample") for l in lines) and any(l.startswith("@PG\t") for l in lines)
    assert body == [l for per_read in case["sam"] for l in per_read]


def test_command_line_bam_output_and_duplicate_names(tmp_path, fake_gpu):
    case = E2E["cases"][0]
    ref, reads = case_inputs(case["name"])
    FakeIndex.contigs = ref
    FakeAligner.rows_by_seq = {seq.upper(): [tuple(r) for r in recs] for (_, seq), recs in zip(reads, case["records"])}
    rpath = _write_inputs(tmp_path, ref, reads + reads, fastq=True)       # every name twice: the second copy is skipped
    out = tmp_path / "out.bam"
    cli.main(["-ref", str(tmp_path / "ref.fa"), "-read", rpath, "-mode", "H", "--nowriteindex", "-o", str(out)])
    _, refs, recs = bam.read_bam_records(str(out))
    want = [l.split("\t") for per_read in case["sam"] for l in per_read]
    assert [(r["name"], r["flag"], refs[r["ref_id"]][0], r["pos"] + 1, r["cigar"]) for r in recs] == \
           [(f[0], int(f[1]), f[2], int(f[3]), f[5]) for f in want]
    assert all(r["qual"] != "*" for r in recs)          # FASTQ qualities are carried (no --Q)


def test_command_line_debug_flag_reports_batches(tmp_path, fake_gpu, capsys):
    case = E2E["cases"][0]
    ref, reads = case_inputs(case["name"])
    FakeIndex.contigs = ref
    FakeAligner.rows_by_seq = {seq.upper(): [tuple(r) for r in recs] for (_, seq), recs in zip(reads, case["records"])}
    rpath = _write_inputs(tmp_path, ref, reads)
    cli.main(["-ref", str(tmp_path / "ref.fa"), "-read", rpath, "-mode", "H", "--nowriteindex", "--debug", "-o", str(tmp_path / "o.sam")])
    err = capsys.readouterr().err
    assert "[vacmap_b200] batch of 1 reads, 3 records" in err and "k_fill 0.5" in err and "n_fill_jobs" not in err


def test_command_line_asm_mode_uses_the_modes_emitter(tmp_path, fake_gpu, monkeypatch):
    """`-mode asm`: contigs one at a time through asm.assembly_align (stubbed: the reference's rows for the 520 kb contig),
    lines from the mode's own emitter -- the reference's asm lines (tests/golden/asm_sam.json.gz, default options + --eqx)."""
    import hashlib
    import json
    import synth
    from vacmap_b200 import asm
    A = json.load(gzip.open(os.path.join(HERE, "golden", "asm_sam.json.gz"), "rt"))
    case = next(c for c in A if not c["variant"]["H"] and not c["variant"]["fakecigar"] and not c["variant"]["md"] and not c["variant"]["qual"])
    ref, read = synth.asm_e2e_inputs()
    FakeIndex.contigs = [(n, s.upper()) for n, s in ref]
    monkeypatch.setattr(asm, "assembly_align", lambda rid, seq, index, opt, ctx=None: [tuple(r) for r in case["records"]])
    rid = case["records"][0][0]
    rpath = _write_inputs(tmp_path, ref, [(rid, read)])
    out = tmp_path / "asm.sam"
    cli.main(["-ref", str(tmp_path / "ref.fa"), "-read", rpath, "-mode", "asm", "-workdir", str(tmp_path), "--nowriteindex", "-o", str(out)])

    def squash(line):
        f = line.split("\t")
        for i in (9, 10):
            if len(f[i]) > 64:
                f[i] = "%d:%s" % (len(f[i]), hashlib.sha1(f[i].encode()).hexdigest())
        return "\t".join(f)
    body = [squash(l) for l in out.read_text().splitlines() if not l.startswith("@")]
    assert body == case["sam"]
