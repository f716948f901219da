"""BAM writer / reader (vacmap_b200/bam.py): the SAM lines the reference's emitter produced (tests/golden/e2e.json.gz)
encoded as BAM and decoded again field by field; unaligned-BAM input as the reference reads it through pysam."""
import gzip
import struct

import pytest

from test_oracle_e2e import E2E, case_inputs
from vacmap_b200 import bam, sam


def _tags_text(aux):
    """BAM aux bytes -> SAM tag strings (integers all print as `i`)."""
    out, p = [], 0
    while p < len(aux):
        tag, typ = aux[p:p + 2].decode(), chr(aux[p + 2])
        p += 3
        if typ == "A":
            out.append("%s:A:%s" % (tag, chr(aux[p]))); p += 1
        elif typ in "cCsSiI":
            fmt = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I"}[typ]
            out.append("%s:i:%d" % (tag, struct.unpack_from(fmt, aux, p)[0])); p += struct.calcsize(fmt)
        elif typ == "f":
            out.append("%s:f:%g" % (tag, struct.unpack_from("<f", aux, p)[0])); p += 4
        elif typ in "ZH":
            e = aux.index(b"\x00", p)
            out.append("%s:%s:%s" % (tag, typ, aux[p:e].decode())); p = e + 1
        elif typ == "B":
            sub = chr(aux[p]); (n,) = struct.unpack_from("<i", aux, p + 1)
            fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[sub]
            vals = struct.unpack_from("<%d%s" % (n, fmt), aux, p + 5)
            out.append("%s:B:%s,%s" % (tag, sub, ",".join(str(v) for v in vals))); p += 5 + n * struct.calcsize(fmt)
        else:
            raise AssertionError(typ)
    return out


@pytest.mark.parametrize("sorted_out", [False, True])
def test_sam_lines_survive_the_bam_round_trip(tmp_path, sorted_out):
    import re
    total = 0
    for ci, case in enumerate(E2E["cases"]):
        ref, _ = case_inputs(case["name"])
        contigs = [(n, len(s)) for n, s in ref]
        lines = [ln for per_read in case["sam"] for ln in per_read]
        total += len(lines)
        header = sam.header_text(contigs, rg={"ID": "1", "SM": "sample"}, command_line="test")
        path = tmp_path / ("o%d.sorted.bam" % ci if sorted_out else "o%d.bam" % ci)
        w = bam.BamWriter(str(path), header)
        w.write_sam_lines(lines)
        w.close()
        raw = path.read_bytes()
        assert raw[:4] == b"\x1f\x8b\x08\x04" and raw[12:16] == b"BC\x02\x00" and raw.endswith(bam._BGZF_EOF)
        text, refs, recs = bam.read_bam_records(str(path))
        assert refs == contigs and ("SO:coordinate" in text) == sorted_out and "@RG\tID:1" in text
        names = [n for n, _ in contigs]
        want = []
        for ln in lines:
            f = ln.split("\t")
            want.append(dict(name=f[0], flag=int(f[1]), ref=f[2], pos=int(f[3]) - 1, mapq=int(f[4]), cigar=f[5], seq=f[9], qual=f[10], tags=f[11:]))
        if sorted_out:
            order = sorted(range(len(want)), key=lambda i: (names.index(want[i]["ref"]), want[i]["pos"], i))
            want = [want[i] for i in order]
        assert len(recs) == len(want)
        for r, w_ in zip(recs, want):
            assert (r["name"], r["flag"], names[r["ref_id"]], r["pos"], r["mapq"], r["cigar"], r["seq"], r["qual"]) == \
                   (w_["name"], w_["flag"], w_["ref"], w_["pos"], w_["mapq"], w_["cigar"], w_["seq"], w_["qual"])
            assert _tags_text(r["aux"]) == w_["tags"]
            ref_len = sum(int(n) for n, op in re.findall(r"(\d+)([MIDNSHP=X])", r["cigar"]) if op in "MDN=X")
            assert r["bin"] == bam.reg2bin(r["pos"], r["pos"] + max(ref_len, 1))
    assert total > 90


def test_reg2bin_known_values():
    # SAM spec 5.3: bin 4681 = first 16 kb leaf; an interval crossing a 16 kb boundary moves one level up
    assert bam.reg2bin(0, 1) == 4681 and bam.reg2bin(16383, 16384) == 4681 and bam.reg2bin(16384, 16385) == 4682
    assert bam.reg2bin(16000, 17000) == 585 and bam.reg2bin(0, 1 << 29) == 0


def test_unaligned_bam_input_matches_pysam_semantics(tmp_path):
    """vacmap:455-466: sequence upper-case on the read's own strand; FLAG 16 records are reverse-complemented and their
    qualities reversed; records without a sequence are skipped; missing qualities give None."""
    header = "@HD\tVN:1.0\tSO:unknown\n"
    lines = ["r1\t4\t*\t0\t0\t*\t*\t0\t0\tACGTN\tIIHG#",
             "r2\t20\t*\t0\t0\t*\t*\t0\t0\tAACG\t*",
             "r3\t4\t*\t0\t0\t*\t*\t0\t0\t*\t*",
             "r4\t16\t*\t0\t0\t*\t*\t0\t0\tAACGR\t!\"#$%"]
    p = tmp_path / "in.bam"
    w = bam.BamWriter(str(p), header)
    w.write_sam_lines(lines)
    w.close()
    got = list(bam.read_bam(str(p)))
    assert got == [("r1", "ACGTN", "IIHG#"), ("r2", "CGTT", None), ("r4", "YCGTT", "%$#\"!")]
    # the command line's reader dispatches on the extension
    from vacmap_b200.__main__ import read_records
    assert list(read_records(str(p))) == got


def test_long_cigar_moves_to_the_cg_tag(tmp_path):
    cigar = "1M1I" * 40000                      # 80 000 operations > 65 535
    seq = "A" * 80000
    line = "q\t0\tchr1\t1\t60\t%s\t*\t0\t0\t%s\t*\tNM:i:40000" % (cigar, seq)
    p = tmp_path / "long.bam"
    w = bam.BamWriter(str(p), "@HD\tVN:1.0\n@SQ\tSN:chr1\tLN:100000\n")
    w.write_sam_lines([line])
    w.close()
    _, _, recs = bam.read_bam_records(str(p))
    assert recs[0]["cigar"] == "80000S40000N"
    tags = _tags_text(recs[0]["aux"])
    assert tags[0] == "NM:i:40000" and tags[1].startswith("CG:B:I,") and tags[1].count(",") == 80000
