"""Shared loader of tests/golden/guide1.npz: anchors the REFERENCE's own get_localmap_multi_all_forDP_inv_guide_1
(mammap_clrnano.py:23069-23345) appended for guide chains of bulk reads (tests/golden/make_bulk.py::gen_guide1)."""
import os

import numpy as np

import bulk

HERE = os.path.dirname(os.path.abspath(__file__))


def jobs():
    """-> list of dict(case, ref, seq (oriented), rc, chain int64[m,4], range (readstart, readend), out int64[n,4])"""
    Z = np.load(os.path.join(HERE, "golden", "guide1.npz"))
    refs, reads = {}, {}
    comp = str.maketrans("ACGTN", "TGCAN")
    res = []
    for j in range(int(Z["n_jobs"])):
        name = str(Z["j%d_case" % j])
        if name not in refs:
            refs[name] = bulk.reference_for(name)
            reads[name] = bulk.reads_for(name, refs[name])
        seq = reads[name][int(Z["j%d_read" % j])][1].upper()
        rc = seq.translate(comp)[::-1]
        if bool(Z["j%d_flip" % j]):
            seq, rc = rc, seq
        res.append(dict(case=name, ref=refs[name], seq=seq, rc=rc, chain=Z["j%d_chain" % j], range=tuple(int(v) for v in Z["j%d_range" % j]),
                        out=Z["j%d_out" % j]))
    return res
