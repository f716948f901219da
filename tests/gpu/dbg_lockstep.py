import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bench
import vacmap_b200 as vb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ref, reads, cat, off = bench.make_workload(n)
ctx = vb._lib.Context(0)
ix = vb.Index(ref, w=10, k=15, ctx=ctx)
al = vb.Aligner(ix, vb.default_option("H"), "H", workers=int(os.environ.get("W", "1")))
for it in range(2):
    rec_off, recs, cig = al.align_packed(cat, off)
    print("records", len(recs), "ops", len(cig), {k: round(v, 1) for k, v in al.last_stage_ms.items() if k.startswith(("k_", "c_", "chain", "total", "extend", "fill", "seed", "reseed", "h_", "g_"))})
