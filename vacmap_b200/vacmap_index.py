"""`vacmap_index`-compatible module over libvacmap_b200.so (the FFI seam).

The reference imports the un-vendored C extension ``vacmap_index`` as ``mp`` and uses exactly this surface:

* ``mp.Aligner(path, w=, k=)`` (``vacmap:344``) with ``.k``, ``.seq_offset`` (``vacmap:358-361``), ``.seq(name)``
  (``vacmap:363``) and ``.map(seq, check_num=, mid_occ=)`` (``mammap_clrnano.py:23985``);
* ``mp.fastx_read(path, read_comment=)`` (``vacmap:445``);
* ``mp.k_cigar(target, query, match, mismatch, gap_open_1, gap_extend_1, gap_open_2, gap_extend_2, bw, zdropvalue[, eqx])``
  at its two call sites: the z-drop edge extension (``2381``: 2,-4,4,4,4,4, bw=100, zdropvalue=50, of which only
  ``q_e`` / ``t_e`` are read, ``2384-2385``) and the global fill (``21554``: 2,-4,4,2,24,1, bw=-1, zdropvalue=-1, eqx).

``sys.modules["vacmap_index"] = vacmap_b200.vacmap_index`` (plus ``edlib_align`` for ``edlib.align``) lets the
reference's own Python run over the CUDA natives call by call.  Every call launches kernels on the GPU: without the
library or a device it raises -- there is no CPU path.  The ``*_batch`` forms take many calls at once (one launch).
"""
import re

import numpy as np

from . import align

fastx_read = align.read_fastx

_EXT = (2, -4, 4, 4, 4, 4, 100, 50)
_FILL = (2, -4, 4, 2, 24, 1, -1, -1)
_CIG = re.compile(r"(\d+)([MIDNSHP=X])")


class Aligner:
    """``vacmap_index.Aligner``: the index lives in HBM (``vm_index_create``); ``map`` runs the seeding kernels."""

    def __init__(self, fn_idx_in=None, w=10, k=15, contigs=None, device=0, **_kw):
        self._ix = align.Index(contigs if contigs is not None else fn_idx_in, w=w, k=k, device=device)
        self.k, self.w = int(k), int(w)

    def __bool__(self):
        return self._ix.h is not None

    @property
    def seq_offset(self):
        return self._ix.seq_offset

    def seq(self, name, start=0, end=0x7fffffff):
        return self._ix.seq(name, start, end)

    def map_batch(self, seqs, check_num=100, mid_occ=-1, arrays=False):
        """``map`` for many reads in one launch -> list of lists of (readpos, refpos_global, strand, len); arrays=True:
        int64 arrays [n, 4] instead of lists of tuples (the contig path's 100 kb slices carry ~20 000 anchors each)."""
        if mid_occ != -1:
            raise NotImplementedError("only the library default occurrence cap (mid_occ=-1) is on the path (clrnano:23985)")
        out = []
        for s, (rows, flipped) in zip(seqs, align.seed_batch(self._ix, [s.upper() for s in seqs], check_num)):
            if flipped:
                # vm_seed_batch_rows hands back the anchors after get_reversed_chain_numpy_rough (:21202-21217);
                # map() itself returns them before: the flip is its own inverse
                rows = rows[::-1].copy()
                rows[:, 0] = len(s) - rows[:, 0] - rows[:, 3]
                rows[:, 2] *= -1
            out.append(np.ascontiguousarray(rows, dtype=np.int64).reshape(-1, 4) if arrays else [tuple(int(v) for v in r) for r in rows])
        return out

    def map(self, seq, check_num=100, mid_occ=-1):
        return self.map_batch([seq], check_num, mid_occ)[0]


def _params(match, mismatch, gap_open_1, gap_extend_1, gap_open_2, gap_extend_2, bw, zdropvalue):
    p = (match, mismatch, gap_open_1, gap_extend_1, gap_open_2, gap_extend_2, bw, zdropvalue)
    if p == _EXT:
        return "extend"
    if p == _FILL:
        return "fill"
    raise NotImplementedError("k_cigar%r: the kernels implement the two parameter sets of the reference's call sites "
                              "(clrnano:2381 and :21554)" % (p,))


def k_cigar_batch(pairs, match=2, mismatch=-4, gap_open_1=4, gap_extend_1=2, gap_open_2=24, gap_extend_2=1, bw=-1,
                  zdropvalue=-1, eqx=False):
    """Many ``k_cigar`` calls with the same parameters: pairs = [(target, query)] -> list of result tuples."""
    kind = _params(match, mismatch, gap_open_1, gap_extend_1, gap_open_2, gap_extend_2, bw, zdropvalue)
    ts = [t.upper() for t, _ in pairs]
    qs = [q.upper() for _, q in pairs]
    if kind == "extend":
        # the path reads q_e / t_e only (:2384-2385, 2413-2414, 2481, 2507): the kernel keeps no traceback
        return [(None, 0, q_e, t_e, 0, 0) for q_e, t_e in align.pairs_batch("extend", ts, qs)]
    out = []
    for t, q, cg in zip(ts, qs, align.pairs_batch("fill", ts, qs, eqx=bool(eqx))):
        ndel = sum(int(n) for n, op in _CIG.findall(cg) if op == "D")
        nins = sum(int(n) for n, op in _CIG.findall(cg) if op == "I")
        out.append((cg, 0, len(q), len(t), ndel, nins))
    return out


def k_cigar(target, query, match=2, mismatch=-4, gap_open_1=4, gap_extend_1=2, gap_open_2=24, gap_extend_2=1, bw=-1,
            zdropvalue=-1, eqx=False):
    """-> (cigar, zdropcode, q_e, t_e, ndel, nins)"""
    return k_cigar_batch([(target, query)], match, mismatch, gap_open_1, gap_extend_1, gap_open_2, gap_extend_2, bw,
                         zdropvalue, eqx)[0]


def edlib_align_batch(pairs):
    """[(query, target)] -> list of {'editDistance': d}"""
    d = align.pairs_batch("distance", [t.upper() for _, t in pairs], [q.upper() for q, _ in pairs])
    return [{"editDistance": int(x)} for x in d]


def edlib_align(query=None, target=None, task="distance", **_kw):
    """``edlib.align(query=, target=, task='distance')`` (clrnano:19251): global NW unit-cost distance."""
    if task != "distance":
        raise NotImplementedError("only task='distance' is on the path (clrnano:19251)")
    return edlib_align_batch([(query, target)])[0]


def install():
    """Put this module in ``sys.modules`` as ``vacmap_index`` and an ``edlib`` stand-in beside it."""
    import sys
    import types
    sys.modules["vacmap_index"] = sys.modules[__name__]
    e = types.ModuleType("edlib")
    e.align = edlib_align
    sys.modules["edlib"] = e
