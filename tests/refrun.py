"""Run the REFERENCE's own per-read driver over the oracle shim (build container only).

This is the end-to-end CPU reference of SURVEY.md 8(c): identical seeding / extension by
construction (the oracle's C restatements of `vacmap_index` and `edlib`), reference
logic for everything else.  Used to generate tests/golden/e2e_*.json.
"""
import numpy as np

import oracle
import oracle.shim as shim
import refimport


def default_option(mode="H", **over):
    skips = {"L": (59., 40., 0.1), "H": (40., 40., 0.2)}.get(mode, (30., 30., 0.5))
    opt = {"mode": mode, "c": 100, "eqx": False, "md": False, "cigar2cg": False, "copycomments": False, "H": False,
           "fakecigar": False, "Q": False, "debug": False, "shortcs": True, "rg-id": "1", "local_kmersize": 9,
           "local_skipcost": skips[0], "golbal_skipcost": skips[1], "maxdivergence": skips[2],
           "golbal_maxdiff": 50, "local_maxdiff": 30, "markunbalancetra": mode in ("L", "H"),
           "nodiscard": mode not in ("L", "H")}
    opt.update(over)
    return opt


_MODE_MODULE = {"H": "clrnano", "L": "ccs", "S": "sensitive", "R": "noprefercloser", "asm": "asm"}


class ReferenceRunner:
    def __init__(self, contigs, mode="H", w=10, k=15, **opt_over):
        shim.install()
        self.mod = refimport.load_mode(_MODE_MODULE[mode])
        self.mod.mp = __import__("vacmap_index")
        self.mod.edlib = __import__("edlib")
        from numba.typed import Dict, List
        from numba import types
        self.aligner = shim.Aligner(contigs=contigs, w=w, k=k)
        self.option = default_option(mode, **opt_over)
        self.contig2start = Dict.empty(types.unicode_type, types.int64)
        self.contig2seq = Dict.empty(types.unicode_type, types.unicode_type)
        self.index2contig = List()
        self.contig2iloc = {}
        for i, item in enumerate(self.aligner.seq_offset):
            name = item[0].decode()
            self.contig2start[name] = item[2]
            self.index2contig.append(name)
            self.contig2iloc[name] = i
            self.contig2seq[name] = self.aligner.seq(name).upper()

    def align(self, readid, seq):
        """-> onemapinfolist (list of 9-tuples) exactly as get_readmap_DP_test returns it; [] on failure."""
        try:
            onemapinfolist, _, _, _ = self.mod.get_readmap_DP_test(
                readid, seq.upper(), self.contig2start, self.contig2seq, self.aligner, self.index2contig,
                self.option, hastra=False, redo_ratio=5, eqx=self.option["eqx"], check_num=self.option["c"])
        except Exception as e:  # the reference worker swallows per-read exceptions (clrnano:24116-24125)
            self.last_error = e
            return []
        return [tuple(r) for r in onemapinfolist]

    def sam_lines(self, readid, seq, qual=None):
        recs = self.align(readid, seq)
        if not recs:
            return []
        o = self.option
        return self.mod.get_bam_dict_str(recs, seq.upper(), qual, self.contig2iloc, self.contig2seq, o["md"],
                                         o["shortcs"], o["cigar2cg"], o["markunbalancetra"], o)
