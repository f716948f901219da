"""Host-glue profile on the CPU-only box: the product driver + glue over the oracle's C stage functions
(tests/gluetest), per-phase wall time of the configs[1] workload.  usage: python tests/glue_profile.py [n_reads] [threads]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth
import test_glue_cpu as tg
import vacmap_b200.align as va

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ref = synth.make_reference(1, 5_000_000)
reads = synth.make_reads(ref, 11, n, read_len=15000, err=0.10)
os.environ["GT_TIMES"] = "1"
L = tg.build_harness()
t0 = time.time()
out = tg.glue_align(L, ref, reads, va.default_option("H"), "H", threads=threads)
print("reads", n, "threads", threads, "wall", round(time.time() - t0, 2), "records", sum(len(o) for o in out))
