"""GPU end-to-end parity: records from the CUDA pipeline (through the C ABI) must equal
(1) the records the REFERENCE's own Python produced (tests/golden/e2e.json.gz) and
(2) the oracle pipeline on fresh seeded reads."""
import numpy as np
import pytest

import oracle
import oracle.pipeline as pl
import synth
from test_oracle_e2e import E2E, case_inputs, option_for

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ci", range(len(E2E["cases"])))
def test_cuda_matches_reference_records(gpu_ctx, ci):
    import vacmap_b200 as vb
    case = E2E["cases"][ci]
    ref, reads = case_inputs(case["name"])
    ix = vb.Index(ref, w=10, k=15, ctx=gpu_ctx)
    al = vb.Aligner(ix, option_for(case), case["mode"])
    got = al.align_batch(reads)
    for (rid, _), g, w in zip(reads, got, case["records"]):
        assert [list(r) for r in g] == w, rid
    ix.close()


def test_index_matches_oracle_index(gpu_ctx):
    import vacmap_b200 as vb
    ref = synth.make_reference(5, 200000, n_contigs=2)
    ix = vb.Index(ref, ctx=gpu_ctx)
    ox = oracle.Index(ref)
    assert (ix.n_keys, ix.n_minimizers, ix.mid_occ) == (ox.n_keys, ox.n_occ, ox.mid_occ_default)
    assert ix.seq("chr2") == ox.seqs[1]
    ix.close()


@pytest.mark.parametrize("seed,mode,kw", [(21, "H", {}), (22, "H", {"eqx": True}), (23, "S", {}), (24, "L", {})])
def test_cuda_matches_oracle_on_fresh_reads(gpu_ctx, seed, mode, kw):
    import vacmap_b200 as vb
    ref = synth.make_reference(seed, 400000, n_contigs=2)
    err = 0.005 if mode == "L" else 0.10
    ratio = (1, 1, 1) if mode == "L" else (4, 3, 3)
    reads = synth.make_reads(ref, seed + 100, 24, read_len=8000, err=err, ratio=ratio, sv_frac=0.4)
    reads.append(("tiny", "ACGT" * 5))
    reads.append(("random", synth.random_seq(np.random.default_rng(1), 3000).tobytes().decode()))
    opt = vb.default_option(mode, **kw)
    ix = vb.Index(ref, ctx=gpu_ctx)
    got = vb.Aligner(ix, opt, mode, workers=1).align_batch(reads)
    ox = oracle.Index(ref)
    ctg = pl.Contigs([n for n, _ in ref], [s for _, s in ref])
    for (rid, seq), g in zip(reads, got):
        want = pl.align_read(rid, seq, ox, ctg, opt, mode)
        assert [tuple(r) for r in g] == [tuple(w) for w in want], rid
    # pipelined sub-batches (own stream / arenas per worker) must give the very same records, in read order
    piped = vb.Aligner(ix, opt, mode, workers=3, chunk_reads=4).align_batch(reads)
    assert piped == got
    al = vb.Aligner(ix, opt, mode, workers=2, chunk_reads=5)
    enc = [s.upper().encode() for _, s in reads]
    off = np.concatenate([[0], np.cumsum([len(e) for e in enc])]).astype(np.int64)
    al.upload_reads(b"".join(enc), off)
    a = al.align_packed(b"".join(enc), off, resident=True)
    b = vb.Aligner(ix, opt, mode, workers=1).align_packed(b"".join(enc), off)
    assert all((x == y).all() for x, y in zip(a, b))
    # two batches in flight at once (submit / wait): same records as one after the other
    al2 = vb.Aligner(ix, opt, mode, workers=2, chunk_reads=7)
    h1 = al2.submit_packed(b"".join(enc), off)
    h2 = al2.submit_packed(b"".join(enc[::-1]), np.concatenate([[0], np.cumsum([len(e) for e in enc[::-1]])]).astype(np.int64))
    r1, r2 = al2.wait(h1), al2.wait(h2)
    assert all((x == y).all() for x, y in zip(r1, b))
    want2 = vb.Aligner(ix, opt, mode, workers=1).align_batch([(rid, s) for rid, s in reads[::-1]])
    n_rec = [len(w) for w in want2]
    assert list(np.diff(r2[0])) == n_rec
    ix.close()


def test_edit_distance_upper_bound_never_undercuts(gpu_ctx, monkeypatch):
    """VM_ED_CHECK makes the backend run the exact kernel next to the bound through the anchors and raise if the
    bound is ever below the distance; the records must still equal the oracle's (modes with 0.2 / 0.5 thresholds)."""
    import vacmap_b200 as vb
    monkeypatch.setenv("VM_ED_CHECK", "1")
    for seed, mode in ((41, "H"), (42, "S")):
        ref = synth.make_reference(seed, 300000, n_contigs=2)
        reads = synth.make_reads(ref, seed + 100, 16, read_len=9000, err=0.12, sv_frac=0.5)
        opt = vb.default_option(mode)
        ix = vb.Index(ref, ctx=gpu_ctx)
        got = vb.Aligner(ix, opt, mode, workers=1).align_batch(reads)
        ox = oracle.Index(ref)
        ctg = pl.Contigs([n for n, _ in ref], [s for _, s in ref])
        for (rid, seq), g in zip(reads, got):
            want = pl.align_read(rid, seq, ox, ctg, opt, mode)
            assert [tuple(r) for r in g] == [tuple(w) for w in want], rid
        ix.close()


@pytest.mark.parametrize("ci", range(len(E2E["cases"])))
def test_cuda_sam_text_matches_reference(gpu_ctx, ci):
    """Reads in, SAM text out (CUDA path + vacmap_b200.sam) == the lines the reference's get_bam_dict_str wrote."""
    import vacmap_b200 as vb
    case = E2E["cases"][ci]
    ref, reads = case_inputs(case["name"])
    ix = vb.Index(ref, w=10, k=15, ctx=gpu_ctx)
    got = vb.Aligner(ix, option_for(case), case["mode"]).sam_lines(reads)
    assert got == case["sam"]
    ix.close()


def test_command_line_on_testdata(tmp_path):
    """BASELINE configs[0]: `-ref testdata/reference.fasta -read testdata/read.fasta -mode H` -> 3 alignments
    (README:124), the inverted middle segment primary (FLAG 16), the flanks supplementary (FLAG 2048)."""
    import gzip
    import os
    import shutil
    import vacmap_b200.__main__ as cli
    td = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "testdata")
    for f in ("reference.fasta", "read.fasta"):
        with gzip.open(os.path.join(td, f + ".gz"), "rb") as fi, open(tmp_path / f, "wb") as fo:
            shutil.copyfileobj(fi, fo)
    out = tmp_path / "out.sam"
    cli.main(["-ref", str(tmp_path / "reference.fasta"), "-read", str(tmp_path / "read.fasta"), "-mode", "H", "-o", str(out)])
    lines = out.read_text().splitlines()
    body = [l for l in lines if not l.startswith("@")]
    assert lines[0] == "@HD\tVN:1.0" and any(l.startswith("@SQ\tSN:") for l in lines)
    assert body == E2E["cases"][0]["sam"][0]
    assert sorted(l.split("\t")[1] for l in body) == ["16", "2048", "2048"]


def test_pinned_host_reads_give_the_same_records(gpu_ctx):
    """vm_host_alloc: a batch packed into page-locked memory (DMA upload per sub-batch) aligns like the same bytes
    from pageable memory, several workers, two jobs in flight."""
    import vacmap_b200 as vb
    ref = synth.make_reference(31, 300000)
    reads = synth.make_reads(ref, 32, 40, read_len=4000, err=0.10, sv_frac=0.3)
    ix = vb.Index(ref, w=10, k=15, ctx=gpu_ctx)
    al = vb.Aligner(ix, vb.default_option("H"), "H", workers=4, chunk_reads=10)
    cat = b"".join(s.encode() for _, s in reads)
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(s) for _, s in reads])
    want = al.align_packed(cat, off)
    pin = gpu_ctx.pinned_bytes(len(cat))
    pin[:] = np.frombuffer(cat, dtype=np.uint8)
    jobs = [al.submit_packed(pin, off) for _ in range(2)]
    for j in jobs:
        got = al.wait(j)
        assert (got[0] == want[0]).all() and (got[1] == want[1]).all() and (got[2] == want[2]).all()
    assert want[0][-1] > 0
    del pin
    ix.close()


def test_command_line_on_two_gpus(tmp_path):
    """One process per GPU (torchrun): index broadcast from rank 0, super-batches round robin, part files stitched in
    input order -- the same SAM body as the single-GPU run."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import vacmap_b200.__main__ as cli
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = synth.make_reference(41, 300000)
    reads = synth.make_reads(ref, 42, 24, read_len=3000, err=0.10, sv_frac=0.3)
    (tmp_path / "ref.fa").write_text("".join(">%s\n%s\n" % (n, s) for n, s in ref))
    (tmp_path / "reads.fa").write_text("".join(">%s\n%s\n" % (n, s) for n, s in reads))
    common = ["-ref", str(tmp_path / "ref.fa"), "-read", str(tmp_path / "reads.fa"), "-mode", "H", "--nowriteindex", "--batch-bases", "9000"]
    cli.main(common + ["-o", str(tmp_path / "one.sam")])
    env = dict(os.environ, PYTHONPATH=root)
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                    "--master-port", "29731", "-m", "vacmap_b200"] + common + ["-o", str(tmp_path / "two.sam")], check=True, env=env, cwd=root,
                   timeout=600)
    body = lambda p: [l for l in p.read_text().splitlines() if not l.startswith("@PG")]
    assert body(tmp_path / "two.sam") == body(tmp_path / "one.sam")
    assert len(body(tmp_path / "one.sam")) > 20
    assert not list(tmp_path.glob("*.part*"))


def test_position_directory_of_long_9mer_runs(gpu_ctx, monkeypatch):
    """A reference with skewed base composition gives 9-mers with hundreds to thousands of occurrences: their position
    runs get a bucket directory (vm_index_build_buckets) and the local re-seeding's window search goes through it.
    Records must equal those of an index built without the directory, and the oracle's."""
    import vacmap_b200 as vb
    rng = np.random.default_rng(77)
    ref = [("chrA", np.frombuffer(b"ACGT", dtype=np.uint8)[rng.choice(4, size=1_500_000, p=[0.46, 0.18, 0.18, 0.18])].tobytes().decode())]
    reads = synth.make_reads(ref, 78, 12, read_len=6000, err=0.10, sv_frac=0.4)
    opt = vb.default_option("H")
    ix = vb.Index(ref, ctx=gpu_ctx)
    got = vb.Aligner(ix, opt, "H", workers=1).align_batch(reads)
    ix.close()
    monkeypatch.setenv("VM_NO_KPOS_DIRECTORY", "1")
    ix0 = vb.Index(ref, ctx=gpu_ctx)
    plain = vb.Aligner(ix0, opt, "H", workers=1).align_batch(reads)
    ix0.close()
    assert got == plain and sum(len(g) for g in got) >= len(reads)
    ox = oracle.Index(ref)
    ctg = pl.Contigs([n for n, _ in ref], [s for _, s in ref])
    for (rid, seq), g in list(zip(reads, got))[:4]:
        want = pl.align_read(rid, seq, ox, ctg, opt, "H")
        assert [tuple(r) for r in g] == [tuple(w) for w in want], rid


def test_command_line_bam_in_bam_out(tmp_path):
    """`-read reads.bam -o out.bam`: the unaligned-BAM input path (vacmap:439-466) and the BAM emitter carry the same three
    alignments as the SAM run on testdata/ (README:124)."""
    import gzip
    import os
    import shutil
    import vacmap_b200.__main__ as cli
    from vacmap_b200 import bam
    td = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "testdata")
    for f in ("reference.fasta", "read.fasta"):
        with gzip.open(os.path.join(td, f + ".gz"), "rb") as fi, open(tmp_path / f, "wb") as fo:
            shutil.copyfileobj(fi, fo)
    import vacmap_b200 as vb
    reads = [(r[0], r[1]) for r in vb.read_fastx(str(tmp_path / "read.fasta"))]
    w = bam.BamWriter(str(tmp_path / "reads.bam"), "@HD\tVN:1.0\tSO:unknown\n")
    w.write_sam_lines(["%s\t4\t*\t0\t0\t*\t*\t0\t0\t%s\t*" % (n, s) for n, s in reads])
    w.close()
    cli.main(["-ref", str(tmp_path / "reference.fasta"), "-read", str(tmp_path / "reads.bam"), "-mode", "H", "--nowriteindex",
              "-o", str(tmp_path / "out.bam")])
    _, refs, recs = bam.read_bam_records(str(tmp_path / "out.bam"))
    want = [l.split("\t") for l in E2E["cases"][0]["sam"][0]]
    assert [(r["name"], r["flag"], refs[r["ref_id"]][0], r["pos"] + 1, r["mapq"], r["cigar"], r["seq"]) for r in recs] == \
           [(f[0], int(f[1]), f[2], int(f[3]), int(f[4]), f[5], f[9]) for f in want]
