"""Score tables of the chaining DPs, built exactly as the reference builds them.

The reference constructs these with numpy at module import
(``mammap_clrnano.py:15371-15376`` ``extra``; ``:26567-26569`` ``readgapcost_list``;
``:27530`` ``log2cache``); numba freezes them as constants.  Building them with the
same numpy expressions on the same host keeps every table entry bit-identical, which
the bit-exact float64 chain scores depend on.
"""
import collections

import numpy as np

ScoreTables = collections.namedtuple("ScoreTables", "extra readgapcost log2cache")
_cached = None


def score_tables():
    global _cached
    if _cached is None:
        extra = []
        g = 0
        while True:
            extra.append(min(36, 30 + 0.5 * np.log(max(g, 1)), min(10, g / 100) + min(30, g / 1000)))
            if len(extra) > 1 and extra[-1] == 36:
                break
            g += 1
        extra = np.ascontiguousarray(np.array(extra, dtype=np.float32))
        readgapcost = np.zeros(100, dtype=np.float32)
        for r in range(1, 100):
            readgapcost[r] = 0.1 * np.log2(r + 1)
        log2cache = np.ascontiguousarray(np.array([0.5 * np.log2(g + 1) for g in range(100000)], dtype=np.float64))
        _cached = ScoreTables(extra, readgapcost, log2cache)
    return _cached
