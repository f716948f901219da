"""Input side of the command line: the FASTA / FASTQ(.gz) reader with the `vacmap_index.fastx_read`
tuple contract (vacmap:445, 474-481) and the read-name de-duplication of the reader loop (vacmap:428-476)."""
import gzip

from vacmap_b200 import align
from vacmap_b200.__main__ import batches


FASTQ = "@r1 XA:i:5\tab:Z:x y\nACGTAC\nGTAC\n+\n@IIIII\nIIII\n@r2\nTTGCA\n+r2\n+@+@+\n\n@r1 again\nAAAA\n+\nIIII\n"
FASTA = ">c1 first contig\nACGT\nacgtn\n\n>c2\nGGGG\n"


def test_fastq_multiline_quality_starting_with_at_and_comments(tmp_path):
    p = tmp_path / "r.fastq"
    p.write_text(FASTQ)
    got = list(align.read_fastx(str(p), read_comment=True))
    assert got == [("r1", "ACGTACGTAC", "@IIIIIIIII", "XA:i:5\tab:Z:x y"), ("r2", "TTGCA", "+@+@+", None),
                   ("r1", "AAAA", "IIII", "again")]
    assert list(align.read_fastx(str(p))) == [r[:3] for r in got]


def test_fasta_gz_multiline_no_quality(tmp_path):
    p = tmp_path / "ref.fa.gz"
    with gzip.open(p, "wt") as f:
        f.write(FASTA)
    assert list(align.read_fastx(str(p))) == [("c1", "ACGTacgtn", None), ("c2", "GGGG", None)]


def test_batches_skip_repeated_names_and_cut_by_bases(tmp_path):
    p1, p2 = tmp_path / "a.fastq", tmp_path / "b.fa"
    p1.write_text(FASTQ)
    p2.write_text(">r2 dup in another file\nCCCC\n>r3\nGGGGGGGG\n")
    bs = list(batches([str(p1), str(p2)], False, 12))
    assert [[r[0] for r in b] for b in bs] == [["r1", "r2"], ["r3"]]          # second r1 and second r2 dropped
    assert bs[0][0][1] == "ACGTACGTAC" and bs[1][0][1] == "GGGGGGGG"
    assert [len(r) for b in list(batches([str(p1)], True, 1 << 30)) for r in b] == [4, 4]


def test_fastq_record_with_empty_sequence_keeps_the_next_record_intact(tmp_path):
    """kseq (behind mp.fastx_read, vacmap:445): a trimmed-to-nothing record is a zero-length read; the record
    after it must come through whole (the header was taken for the quality string)."""
    p = tmp_path / "e.fastq"
    p.write_text("@r1\n\n+\n\n@r2\nACGT\n+\nIIII\n@r3\n+\n@r4\nGG\n+\n@>\n")
    assert list(align.read_fastx(str(p))) == [("r1", "", None), ("r2", "ACGT", "IIII"), ("r3", "", None), ("r4", "GG", "@>")]


def test_stitch_parts_restores_batch_order(tmp_path):
    """Multi-GPU command line: batch b is the (b // world)-th block of rank (b % world)'s part file."""
    import io
    from vacmap_b200.__main__ import stitch_parts
    world = 3
    blocks = [("batch%d\n" % b) * (b % 4) for b in range(10)]       # some batches write nothing
    paths, sizes = [], []
    for r in range(world):
        mine = [blocks[b].encode() for b in range(r, 10, world)]
        p = tmp_path / ("o.part%d" % r)
        p.write_bytes(b"".join(mine))
        paths.append(str(p))
        sizes.append([len(m) for m in mine])
    out = io.BytesIO()
    stitch_parts(out, paths, sizes)
    assert out.getvalue().decode() == "".join(blocks)
