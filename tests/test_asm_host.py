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
