"""The library's SAM emitter (csrc/vm_sam.cu, host threads, no device) against the Python emitter sam.get_bam_dict_str --
itself byte-identical to the reference's lines (tests/test_sam.py) -- and against the reference's own lines directly."""
import gzip
import json
import os

import numpy as np
import pytest

from test_oracle_e2e import E2E, case_inputs, option_for
from vacmap_b200 import sam
from vacmap_b200.align import RECORD_DTYPE, OPS

HERE = os.path.dirname(os.path.abspath(__file__))
_ENC = {c: i for i, c in enumerate(OPS)}


def pack_records(per_read_rows, contig_names):
    """rows (readid, contig, strand, q_st, q_en, r_st, r_en, mapq, cigar string) per read -> (rec_off, recs, cig)"""
    return sam.pack_rows(per_read_rows, contig_names)


def python_lines(rows, seq, qual, comment, c2i, c2s, opt, md, shortcs, cigar2cg, mark, copycomments):
    try:
        if copycomments:
            return sam.get_bam_dict_str_comments(rows, seq, qual, comment, c2i, c2s, md, shortcs, cigar2cg, mark, opt)
        return sam.get_bam_dict_str(rows, seq, qual, c2i, c2s, md, shortcs, cigar2cg, mark, opt)
    except Exception:
        return []          # the command line skips the read, as the reference's worker does


def test_native_emitter_reproduces_the_references_lines():
    total = 0
    for case in E2E["cases"]:
        ref, reads = case_inputs(case["name"])
        opt = option_for(case)
        names = [n for n, _ in ref]
        table = sam.ContigTable(ref)
        rows = [[tuple(r) for r in recs] for recs in case["records"]]
        rec_off, recs, cig = pack_records(rows, names)
        batch = [(rid, seq.upper()) for rid, seq in reads]
        data, off = sam.batch_text(batch, rec_off, recs, cig, table, opt, md=opt.get("md", False), shortcs=opt.get("shortcs", True),
                                   cigar2cg=opt.get("cigar2cg", False), markunbalancetra=opt.get("markunbalancetra", False))
        for i, want in enumerate(case["sam"]):
            got = data[off[i]:off[i + 1]].decode().splitlines()
            assert got == want, (case["name"], reads[i][0])
            total += len(want)
    assert total > 90


def _random_cigar(rng, qspan, eqx, clip_front, clip_back, clipsyb):
    """A CIGAR consuming exactly `qspan` query bases; returns (cigar, reference span)."""
    ops, q, r = [], 0, 0
    while q < qspan:
        n = int(min(qspan - q, rng.integers(1, 60)))
        kind = rng.choice(["m", "m", "m", "x", "i", "d"]) if eqx else rng.choice(["M", "M", "M", "i", "d"])
        if kind in ("m", "M"):
            ops.append("%d%s" % (n, "=" if eqx else "M")); q += n; r += n
        elif kind == "x":
            n = min(n, 3); ops.append("%dX" % n); q += n; r += n
        elif kind == "i":
            n = min(n, 8); ops.append("%dI" % n); q += n
        else:
            n = int(rng.integers(1, 9)); ops.append("%dD" % n); r += n
    # adjacent equal ops on purpose now and then (mergecigar_ has to merge them)
    if rng.random() < 0.3:
        ops.insert(1, ops[0])
        n0 = int(ops[0][:-1]); op0 = ops[0][-1]
        q += n0 if op0 in "M=XI" else 0
        r += n0 if op0 in "M=XD" else 0
    front = "%d%s" % (clip_front, clipsyb) if clip_front > 0 else ""
    back = "%d%s" % (clip_back, clipsyb) if clip_back > 0 else ""
    return front + "".join(ops) + back, r, q


@pytest.mark.parametrize("seed", range(6))
def test_native_emitter_equals_python_emitter_on_random_records(seed):
    """Random multi-record reads through every option combination: --MD / cs short and long, --H, --fakecigar, --L,
    --markunbalancetra, qualities, copied comments; records whose walk runs off a sequence (the reference raises) included."""
    rng = np.random.default_rng(seed)
    bases = np.frombuffer(b"ACGT", np.uint8)
    contigs = [("chr%d" % (i + 1), bases[rng.integers(0, 4, size=4000)].tobytes().decode()) for i in range(3)]
    contigs[1] = (contigs[1][0], contigs[1][1][:1500] + "N" * 40 + contigs[1][1][1540:])
    c2s = dict(contigs)
    c2i = {n: i for i, (n, _) in enumerate(contigs)}
    table = sam.ContigTable(contigs)
    for combo in range(12):
        md, shortcs = bool(combo & 1), bool(combo & 2)
        hard, fake = bool(combo & 4), bool(combo & 8)
        cigar2cg, mark, copyc = combo % 3 == 0, combo % 4 == 1, combo % 2 == 0
        eqx = md or combo % 5 == 0
        opt = {"H": hard, "fakecigar": fake, "rg-id": "grp%d" % combo}
        clipsyb = "H" if hard else "S"
        reads, rows_all = [], []
        for ri in range(25):
            qlen = int(rng.integers(200, 900))
            seq = bases[rng.integers(0, 4, size=qlen)].tobytes().decode()
            if ri % 7 == 3:
                seq = seq[:50] + "RYKM" + seq[54:]          # IUPAC codes in a read
            qual = None if ri % 3 == 0 else "".join(chr(33 + int(v)) for v in rng.integers(0, 60, size=qlen if ri % 11 else qlen - 1))
            comment = [None, "zm:i:5\tRG:Z:other\tbad\txx:Q:1\tqs:Z:a:b", "NM:i:3\tab:f:1.5\tab:Z:dup", ""][ri % 4]
            rows = []
            for k in range(int(rng.integers(0, 5))):
                q_st = int(rng.integers(0, qlen // 2))
                q_en = int(rng.integers(q_st + 20, qlen + 1))
                cname = contigs[int(rng.integers(0, 3))][0]
                cigar, rspan, qcons = _random_cigar(rng, q_en - q_st, eqx, q_st, qlen - q_en, clipsyb)
                r_st = int(rng.integers(0, 4000 - rspan - 1)) if rspan < 3900 else 0
                if ri % 13 == 5 and k == 0:
                    r_st = 4000 - rspan // 2                # runs off the contig: the reference raises, the read is skipped
                rows.append((("read%d" % ri), cname, "+" if rng.random() < 0.5 else "-", q_st, q_en, r_st, r_st + rspan,
                             int(rng.integers(0, 61)), cigar))
            reads.append(("read%d" % ri, seq, qual, comment))
            rows_all.append(rows)
        rec_off, recs, cig = pack_records(rows_all, [n for n, _ in contigs])
        data, off = sam.batch_text(reads, rec_off, recs, cig, table, opt, md=md, shortcs=shortcs, cigar2cg=cigar2cg,
                                   markunbalancetra=mark, copycomments=copyc, threads=3)
        n_lines = 0
        for i, (rid, seq, qual, comment) in enumerate(reads):
            want = python_lines([tuple(r) for r in rows_all[i]], seq, qual, comment, c2i, c2s, opt, md, shortcs, cigar2cg, mark, copyc) \
                if rows_all[i] else []
            got = data[off[i]:off[i + 1]].decode().splitlines()
            assert got == want, (seed, combo, rid)
            n_lines += len(want)
        assert n_lines > 10


def test_native_emitter_long_cigar_goes_to_cg():
    contigs = [("chr1", "AC" * 40000)]
    table = sam.ContigTable(contigs)
    cigar = "1M1I" * 33000                       # 66 000 operations -> 132 000 entries of the reference's oplist
    seq = "A" * 66000
    rows = [[("q", "chr1", "+", 0, 66000, 0, 33000, 60, cigar)]]
    rec_off, recs, cig = pack_records(rows, ["chr1"])
    opt = {"H": False, "fakecigar": False, "rg-id": "1"}
    for l_flag in (True, False):
        data, off = sam.batch_text([("q", seq)], rec_off, recs, cig, table, opt, cigar2cg=l_flag)
        want = sam.get_bam_dict_str([tuple(rows[0][0])], seq, None, {"chr1": 0}, dict(contigs), False, True, l_flag, False, opt)
        assert data.decode().splitlines() == want
        assert ("\tCG:Z:" in data.decode()) == l_flag


def test_native_emitter_asm_mode_matches_reference():
    """asm mode's emitter in the library (`asm_mode`): the reference's own lines for the asm end-to-end records
    (tests/golden/asm_sam.json.gz): NM from the CIGAR alone, the primary-record rule, MAPQ 60 / 1."""
    import hashlib
    import synth

    def squash(line):
        f = line.split("\t")
        for i in (9, 10):
            if len(f[i]) > 64:
                f[i] = "%d:%s" % (len(f[i]), hashlib.sha1(f[i].encode()).hexdigest())
        return "\t".join(f)
    A = json.load(gzip.open(os.path.join(HERE, "golden", "asm_sam.json.gz"), "rt"))
    ref, read = synth.asm_e2e_inputs()
    table = sam.ContigTable([(n, s.upper()) for n, s in ref])
    qual = "".join(chr(33 + (i * 11) % 41) for i in range(len(read)))
    for case in A:
        v = case["variant"]
        opt = {"H": v["H"], "fakecigar": v["fakecigar"], "rg-id": "1"}
        rows = [[tuple(r) for r in case["records"]]]
        rec_off, recs, cig = pack_records(rows, [n for n, _ in ref])
        data, off = sam.batch_text([(rows[0][0][0], read.upper(), qual if v["qual"] else None)], rec_off, recs, cig, table, opt, md=v["md"],
                                   shortcs=v.get("shortcs", True), asm=True)
        assert [squash(x) for x in data.decode().splitlines()] == case["sam"], v
