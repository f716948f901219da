"""GPU box: time the device index build for references of growing size (SURVEY 8f-3 / configs[2])."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import vacmap_b200 as vb
ctx = vb._lib.Context(0)
rng = np.random.default_rng(2)
for mb, k in [(5, 15), (250, 15), (int(sys.argv[1]) if len(sys.argv) > 1 else 1000, 19)]:
    n = mb * 1_000_000
    per = 50_000_000
    t0 = time.time()
    ref = []
    B = np.frombuffer(b"ACGT", dtype=np.uint8)
    for c in range((n + per - 1) // per):
        ln = min(per, n - c * per)
        ref.append(("chr%d" % (c + 1), B[rng.integers(0, 4, size=ln, dtype=np.uint8)].tobytes()))
    t1 = time.time()
    ix = vb.Index(ref, w=10, k=k, ctx=ctx)
    t2 = time.time()
    print("ref %d Mb k%d: generate %.1f s, Index() %.2f s (python marshalling + device build), keys %d minimizers %d mid_occ %d" %
          (mb, k, t1 - t0, t2 - t1, ix.n_keys, ix.n_minimizers, ix.mid_occ), flush=True)
    ix.close()
