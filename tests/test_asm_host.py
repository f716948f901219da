"""Host logic of vacmap_b200.asm (batch loop, carry slice, traceback, overlap trimming) on a box without a GPU: the
DP call is replaced by the oracle's, everything else is the product code; expected paths = the reference-pinned
golden flows (tests/golden/asm_linked.npz)."""
import os

import numpy as np
import pytest

import oracle
from vacmap_b200 import asm
from vacmap_b200.chain import ChainParams

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "asm_linked.npz"))


def _first(gs, gi, pS, pP, prl, a):
    g, S, P, A, _ = oracle.chain_linked_d_all(gs, gi, pS, pP, prl, a, 15, 40., 50, 1000)
    if g == -1:
        g, S, P, A = oracle.chain_linked_fast(gs, gi, pS, pP, prl, a, 15, 40., 50, 1000)
    return g, S, P, A


def _second(gs, gi, pS, pP, prl, a):
    g, S, P, A, _ = oracle.chain_linked_d_all(gs, gi, pS, pP, prl, a, 9, 30., 30, 99, local=True)
    return g, S, P, A


def test_first_round_paths():
    prm = ChainParams(kmersize=15, skipcost=40.0, maxdiff=50, maxgap=1000)
    for fi in range(int(G["n_flows"])):
        batches = [G["f%d_b%d" % (fi, bi)].astype(np.int64).reshape(-1, 4) for bi in range(int(G["f%d_nb" % fi]))]
        path = asm.linked_chain_path(batches, prm, dp=_first)
        assert np.array_equal(np.array(path, dtype=np.int64).reshape(-1, 4), G["f%d_path" % fi]), fi
    assert asm.linked_chain_path([], prm, dp=_first) == []


def test_second_round_paths_trimmed_and_the_traceback_quirk():
    prm = ChainParams(kmersize=9, skipcost=30.0, maxdiff=30, maxgap=99)
    n_err = 0
    for fi in range(int(G["n_lflows"])):
        batches = [G["l%d_b%d" % (fi, bi)].astype(np.int64).reshape(-1, 4) for bi in range(int(G["l%d_nb" % fi]))]
        if int(G["l%d_err" % fi]):
            with pytest.raises(IndexError):
                asm.linked_chain_path(batches, prm, second_round=True, dp=_second)
            n_err += 1
            continue
        path = asm.trim_overlaps(asm.linked_chain_path(batches, prm, second_round=True, dp=_second))
        assert np.array_equal(np.array(path, dtype=np.int64).reshape(-1, 4), G["l%d_path" % fi]), fi
    assert n_err == 1


def test_product_link_cigar_matches_reference():
    """vacmap_b200.asm.link_cigar (boundary-only) and its linear-time fold against the reference's link_cigar
    (mammap_asm.py:22366-22410) on the recorded pairs (tests/golden/asm_link_cigar.json)."""
    import json
    import os
    from vacmap_b200 import asm
    rows = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "asm_link_cigar.json")))
    assert len(rows) >= 300
    for a, b, want in rows:
        assert asm.link_cigar(a, b) == want, (a, b)
    # the fold: joining many pieces one boundary at a time equals the pairwise chain
    import random
    rnd = random.Random(3)
    for _ in range(50):
        pieces = ["".join("%d%s" % (rnd.randint(1, 30), rnd.choice("MID=X")) for _ in range(rnd.randint(1, 4))) for _ in range(rnd.randint(1, 12))]
        ref = pieces[0]
        for p in pieces[1:]:
            ref = asm.link_cigar(ref, p)
        assert asm.link_cigars(pieces) == ref


class _OracleNatives:
    """Stand-ins for the CUDA entry points the asm product path calls, made of the oracle's C natives: the product's own
    host loop (vacmap_b200/asm.py) then runs end to end on a box without a GPU.  Test infrastructure only."""

    def __init__(self, ref, k=15, w=10):
        import oracle.pipeline as pl
        self.ox = oracle.Index(ref, w=w, k=k)
        self.ctg = pl.Contigs([n for n, _ in ref], [s for _, s in ref])
        self.k, self.w = k, w
        self.names = [n for n, _ in ref]
        self.starts = list(self.ctg.starts)
        self.ctx = None
        self._seqs = dict(ref)

    def seq(self, name, start=0, end=0x7fffffff):
        return self._seqs[name][start:end]

    # Aligner.map_batch(..., arrays=True)
    def map_batch(self, seqs, check_num=100, mid_occ=-1, arrays=False):
        out = []
        for s in seqs:
            rows = np.array(self.ox.map(s, check_num=check_num, mid_occ=mid_occ), dtype=np.int64).reshape(-1, 4)
            out.append(rows if arrays else [tuple(int(v) for v in r) for r in rows])
        return out

    def chain_linked_batch(self, jobs, params, ctx=None):
        from vacmap_b200.chain import LinkedChainResult
        res = []
        for gs, gi, pS, pP, prl, a in jobs:
            if params.variant == 4:
                g, S, P, A, _ = oracle.chain_linked_d_all(gs, gi, pS, pP, prl, a, params.kmersize, params.skipcost, params.maxdiff,
                                                          params.maxgap, local=True)
                res.append(LinkedChainResult(g, S, P, A, 0))
                continue
            g, S, P, A, _ = oracle.chain_linked_d_all(gs, gi, pS, pP, prl, a, params.kmersize, params.skipcost, params.maxdiff,
                                                      params.maxgap)
            uf = 0
            if g < 0:          # opcount bail-out: the caller's heuristic twin (mammap_asm.py:23246-23247)
                g, S, P, A = oracle.chain_linked_fast(gs, gi, pS, pP, prl, a, params.kmersize, params.skipcost, params.maxdiff,
                                                      params.maxgap)[:4]
                uf = 1
            res.append(LinkedChainResult(g, S, P, A, uf))
        return res

    def local_reseed_batch(self, index, reads, jobs):
        from vacmap_b200.sam import reverse_complement
        out = []
        for ri, wins, guides, rs, re_ in jobs:
            seq = reads[ri]
            out.append(oracle.local_reseed_scan(self.ctg, wins, np.asarray(guides, dtype=np.int64), seq, reverse_complement(seq), 9, rs, re_))
        return out

    @staticmethod
    def k_cigar(target, query, *a, **kw):
        return oracle.k_cigar(target, query, *a, **kw)

    @staticmethod
    def k_cigar_batch(pairs, *a):
        return [oracle.k_cigar(t, q, *a) for t, q in pairs]


@pytest.mark.parametrize("fixture", ["asm_e2e", "asm_e2e2"])
def test_product_contig_path_over_oracle_natives_matches_reference(monkeypatch, fixture):
    """vacmap_b200.asm.assembly_align -- the PRODUCT's host loop -- with every CUDA entry point swapped for the oracle's
    native of the same contract gives the rows and CIGARs the reference's assembly_get_readmap_DP_test gave for the 520 kb
    contig read (tests/golden/asm_e2e.json.gz) and for the reverse-strand contig with a translocated piece, a tandem
    duplication and a deletion (asm_e2e2.json.gz).  The GPU twin of the first case (tests/test_zz_gpu_asm_host.py) runs the
    same loop over the real entry points."""
    import gzip
    import json
    import synth
    import vacmap_b200 as vb
    from vacmap_b200 import align
    here = os.path.dirname(os.path.abspath(__file__))
    E = json.load(gzip.open(os.path.join(here, "golden", fixture + ".json.gz"), "rt"))
    ref, read = synth.asm_e2e_inputs() if fixture == "asm_e2e" else synth.asm_e2e_inputs_2()
    rid = "ctgread" if fixture == "asm_e2e" else "ctgread2"
    nat = _OracleNatives(ref)
    monkeypatch.setattr(asm, "chain_linked_batch", nat.chain_linked_batch)
    monkeypatch.setattr(align, "local_reseed_batch", nat.local_reseed_batch)
    monkeypatch.setattr(asm._vi, "k_cigar", nat.k_cigar)
    monkeypatch.setattr(asm._vi, "k_cigar_batch", nat.k_cigar_batch)
    monkeypatch.setattr(asm._vi.Aligner, "map_batch", lambda self, seqs, check_num=100, mid_occ=-1, arrays=False:
                        nat.map_batch(seqs, check_num, mid_occ, arrays))
    for case in E["cases"]:
        opt = vb.default_option("S", eqx=case["eqx"])
        opt.update({"golbal_skipcost": 30., "golbal_maxdiff": 50, "local_skipcost": 30., "local_maxdiff": 30, "local_kmersize": 9})
        got = asm.assembly_align(rid, read, nat, opt)
        assert [list(r) for r in got] == case["records"], case["eqx"]
