"""Stage golden of local re-seeding: the oracle's guide_1 restatement (window construction + orc_reseed.c
scan and same-diagonal merge) against the anchors the REFERENCE's own njit function produced (tests/golden/guide1.npz):
same anchors, same emission order, same (readstart, readend)."""
import numpy as np

import guide1_cases
import oracle.pipeline as pl


def test_oracle_local_reseed_matches_reference_guide_1():
    jobs = guide1_cases.jobs()
    assert len(jobs) >= 30 and sum(len(j["out"]) for j in jobs) > 10000
    ctgs = {}
    for j in jobs:
        if j["case"] not in ctgs:
            ctgs[j["case"]] = pl.Contigs([n for n, _ in j["ref"]], [s for _, s in j["ref"]])
        ctg = ctgs[j["case"]]
        out = []
        pl.local_reseed(out, j["chain"], j["seq"], j["rc"], ctg, 9)
        got = np.array(out, dtype=np.int64).reshape(-1, 4)
        assert got.shape == j["out"].shape and (got == j["out"]).all()
        wins, raw = pl.guide_windows(j["chain"], ctg)
        L = len(j["seq"])
        assert (max(0, int(raw[0][0]) - 7000), min(L - 9 + 1, int(raw[-1][0]) + 7000)) == j["range"]
