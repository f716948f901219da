"""ORACLE shim: a `vacmap_index` / `edlib` stand-in over the C restatements.

Installing these into ``sys.modules`` lets the reference's OWN Python
(`get_readmap_DP_test`, `get_bam_dict_str`, ...) run end-to-end in the build container,
which is how the end-to-end golden fixtures are produced (tests/golden/make_golden.py).
Test infrastructure only.
"""
import sys
import types

import oracle


def read_fastx(path, read_comment=False):
    """Minimal FASTA/FASTQ reader with the mappy.fastx_read tuple contract."""
    import gzip
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        name, comment, seq, qual, mode = None, None, [], [], None
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            if line[0] in ">@" and mode != "qual":
                if name is not None:
                    yield _rec(name, comment, seq, qual, read_comment)
                hdr = line[1:].split(None, 1)
                name, comment = hdr[0], (hdr[1] if len(hdr) > 1 else None)
                seq, qual, mode = [], [], ("fa" if line[0] == ">" else "fq")
            elif line[0] == "+" and mode == "fq":
                mode = "qual"
            elif mode == "qual":
                qual.append(line)
                if sum(map(len, qual)) >= sum(map(len, seq)):
                    mode = "fq_done"
            else:
                seq.append(line)
        if name is not None:
            yield _rec(name, comment, seq, qual, read_comment)


def _rec(name, comment, seq, qual, read_comment):
    s = "".join(seq)
    q = "".join(qual) if qual else None
    return (name, s, q, comment) if read_comment else (name, s, q)


class Aligner:
    """vacmap_index.Aligner surface used by the reference (vacmap:344,358-367; clrnano:23985,24024)."""

    def __init__(self, fn_idx_in=None, w=10, k=15, contigs=None, **kw):
        if contigs is None:
            contigs = [(n, s) for n, s, _ in read_fastx(fn_idx_in)]
        self._ix = oracle.Index(contigs, w=w, k=k)
        self.k, self.w = k, w

    @property
    def seq_offset(self):
        return [(n.encode(), len(s), int(o)) for n, s, o in zip(self._ix.names, self._ix.seqs, self._ix.offsets[:-1])]

    def seq(self, name, start=0, end=0x7fffffff):
        return self._ix.seqs[self._ix.names.index(name)][start:end]

    def map(self, seq, check_num=100, mid_occ=-1):
        return [tuple(int(v) for v in r) for r in self._ix.map(seq, check_num, mid_occ)]


class _Edlib:
    @staticmethod
    def align(query, target, task="distance", **kw):
        return {"editDistance": oracle.edit_distance(query, target)}


def install():
    """Put the shim modules in sys.modules (before importing the reference's mode module)."""
    m = types.ModuleType("vacmap_index")
    m.Aligner = Aligner
    m.k_cigar = oracle.k_cigar
    m.fastx_read = read_fastx
    sys.modules["vacmap_index"] = m
    e = types.ModuleType("edlib")
    e.align = _Edlib.align
    sys.modules["edlib"] = e
    return m
