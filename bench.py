#!/usr/bin/env python
"""bench.py -- aligned Gbp/s of the vacmap_b200 per-read alignment path on N B200s.

Contract: `python bench.py --gpus N --steps K --warmup W` (under torchrun for N > 1, one rank
per GPU) prints ONE JSON line on rank 0.  A step = one pass of the whole hot path (seeding ->
global chaining -> local re-seeding + chaining -> extension -> records) over one batch of
synthetic reads per GPU.  `--impl reference` times the CPU oracle port of the same path on
the host cores (the reference is Python + numba over an un-vendored C extension that cannot
be built here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

READ_LEN = 15000
ERR = 0.10
REF_LEN = 5_000_000
WORKLOAD = "10k synthetic ONT reads (15 kb, 10% err) vs 5 Mb reference, -mode H, 1xB200 (BASELINE configs[1])"

# The default run is BASELINE configs[1] (the configuration the metric is quoted on).  --workload cfg2 / cfg3 run the
# other single-GPU-sized configurations per GPU (their bench lines are kept under profiles/, they are not the headline).
WORKLOADS = {
    "cfg1": dict(name=WORKLOAD, mode="H", k=15, w=10, ref_len=REF_LEN, ref_seed=1, n_contigs=1, read_len=READ_LEN, err=ERR, ratio=(4, 3, 3),
                 reads=10000, read_seed=11, sv=False),
    "cfg2": dict(name="synthetic HiFi reads (20 kb, 0.5% err) vs GRCh38-sized (3.1 Gb) reference, -mode L -k 19 -w 10 "
                      "(BASELINE configs[2]; 7 500 reads = 150 Mbp per GPU per step)",
                 mode="L", k=19, w=10, ref_len=3_100_000_000, ref_seed=2, n_contigs=62, read_len=20000, err=0.005, ratio=(1, 1, 1),
                 reads=7500, read_seed=12, sv=False, workers=4, ahead=2),      # six workers' arenas + the 33 GB index + the lock-step leg do not fit
    "cfg3": dict(name="10 kb reads (10% err) from donors with nested DEL/INS/INV/DUP/TRA events (vacsim grammar) vs 250 Mb "
                      "reference, -mode S (BASELINE configs[3]; 10 000 reads per GPU per step)",
                 mode="S", k=15, w=10, ref_len=250_000_000, ref_seed=3, n_contigs=5, read_len=10000, err=0.10, ratio=(4, 3, 3),
                 reads=10000, read_seed=31, sv=True),
    "cfg4": dict(name="synthetic contigs (inversion + deletion + insertion, 0.1% divergence) vs GRCh38-sized (3.1 Gb) reference, "
                      "-mode asm --H --fakecigar (BASELINE configs[4] at reduced contig length; one contig per GPU per step)",
                 mode="asm", k=15, w=10, ref_len=3_100_000_000, ref_seed=2, n_contigs=62, read_len=2_000_000, err=0.001, ratio=(1, 1, 1),
                 reads=1, read_seed=41, sv=False),
}
WL = WORKLOADS["cfg1"]


_STDOUT = None


def emit(line):
    """The one JSON line, on the real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_STDOUT, data)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_indices):
        # one sampler process for all the job's GPUs (rank 0 runs it): eight nvidia-smi loops polling the driver at once
        # are a measurable disturbance on the 4-vCPU-per-rank boxes
        self.gpu, self.rows, self.proc = ",".join(str(g) for g in gpu_indices), [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()

        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = [num(r[1]) for r in self.rows if len(r) >= 9 and num(r[1]) is not None]
        mx = [num(r[2]) for r in self.rows if len(r) >= 9 and num(r[2]) is not None]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        per_gpu = {}
        for r in self.rows:
            if len(r) >= 9 and num(r[1]) is not None:
                per_gpu.setdefault(r[0], []).append(num(r[1]))
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "sm_mhz_per_gpu": {g: float(np.median(v)) for g, v in sorted(per_gpu.items())}}


def make_workload(n_reads, rank=0, ref_only=False):
    """The workload's synthetic inputs: i.i.d. reference with 5 % diverged repeats in 50 Mb contigs; reads with
    i.i.d. errors, seed + rank; cfg3: every read's donor segment carries 1-20 adjacent events of vacsim's grammar."""
    import synth
    ref = synth.make_reference(WL["ref_seed"], WL["ref_len"], n_contigs=WL["n_contigs"])
    if ref_only:
        return ref
    if WL["sv"]:
        import bulk
        rng = np.random.default_rng(WL["read_seed"] + rank)
        arrs = [np.frombuffer(s.encode(), dtype=np.uint8) for _, s in ref]
        reads = []
        for i in range(n_reads):
            a = arrs[int(rng.integers(0, len(arrs)))]
            st = int(rng.integers(0, len(a) - WL["read_len"]))
            seg = bulk.nested_sv(rng, a[st:st + WL["read_len"]].copy(), a, int(rng.integers(1, 21)))
            if rng.random() < 0.5:
                seg = synth._COMP[seg][::-1]
            reads.append(("read_%d" % i, synth.mutate(rng, seg, WL["err"], WL["ratio"]).tobytes().decode()))
    else:
        reads = synth.make_reads(ref, WL["read_seed"] + rank, n_reads, read_len=WL["read_len"], err=WL["err"], ratio=WL["ratio"])
    enc = [s.encode() for _, s in reads]
    off = np.zeros(n_reads + 1, dtype=np.int64)
    for i, e in enumerate(enc):
        off[i + 1] = off[i] + len(e)
    return ref, reads, b"".join(enc), off


# ---------------------------------------------------------------------------
# CPU arm: the oracle port of the whole path, one process per core
# ---------------------------------------------------------------------------
_CPU = {}


def _cpu_init(ref):
    import oracle
    import oracle.pipeline as pl
    oracle.tables()
    _CPU["ix"] = oracle.Index(ref, w=WL["w"], k=WL["k"])
    _CPU["ctg"] = pl.Contigs([n for n, _ in ref], [s for _, s in ref])


def _cpu_work(chunk):
    import oracle.pipeline as pl
    import vacmap_b200.align as va
    opt = va.default_option(WL["mode"])
    bases = 0
    for rid, seq in chunk:
        if pl.align_read(rid, seq, _CPU["ix"], _CPU["ctg"], opt, WL["mode"]):
            bases += len(seq)
    return bases


class CpuArm:
    """The oracle port of the whole path on `cores` worker processes that live for the whole measurement (index built
    once and inherited by fork, every worker warmed before anything is timed)."""

    def __init__(self, ref, cores):
        import multiprocessing as mp
        self.cores = cores
        _cpu_init(ref)                       # built once, inherited by the forked workers
        self.pool = mp.get_context("fork").Pool(cores) if cores > 1 else None

    def run(self, reads):
        """-> (aligned Gbp/s, seconds) over `reads`, cut into many small tasks so the tail stays balanced"""
        step = max(1, len(reads) // (self.cores * 8))
        chunks = [reads[i:i + step] for i in range(0, len(reads), step)]
        t0 = time.perf_counter()
        if self.pool is None:
            bases = sum(_cpu_work(c) for c in chunks)
        else:
            bases = sum(self.pool.imap_unordered(_cpu_work, chunks))
        dt = time.perf_counter() - t0
        return bases / dt / 1e9, dt

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def cpu_inputs(sample, ref=None):
    """Reference and reads of the bounded CPU sample.  For the GRCh38-sized workload the oracle's single-threaded index
    build would take many minutes, so its sample is drawn from (and indexed over) the first 250 Mb of the same reference;
    the per-read cost (chaining, extension) does not depend on the rest."""
    import synth
    if ref is None:
        ref = make_workload(0, ref_only=True)
    note = ""
    if WL["ref_len"] > 1_000_000_000:
        ref = ref[:5]
        note = "; reference cut to its first 250 Mb for the CPU arm"
    reads = synth.make_reads(ref, WL["read_seed"], sample, read_len=WL["read_len"], err=WL["err"], ratio=WL["ratio"]) if not WL["sv"] \
        else make_workload(sample)[1]
    return ref, reads, note


def _cpulist(text):
    out = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        out.extend(range(int(lo), int(hi or lo) + 1))
    return out


def pin_to_gpu_node(local_rank, local_world):
    """Bind this rank to its slice of the cores of the NUMA node its GPU is attached to (sysfs; no-op when the
    topology cannot be read).  Returns a one-line description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        nodes = []
        for i in range(local_world):
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(i)).busId
            bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
            if len(bus.split(":")[0]) == 8:
                bus = bus[4:]
            try:
                nodes.append(int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read()))
            except OSError:
                nodes.append(-1)
        node = nodes[local_rank]
        allowed = sorted(os.sched_getaffinity(0))
        if node < 0:
            cpus, peers = allowed, list(range(local_world))
        else:
            cpus = [c for c in _cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read()) if c in set(allowed)]
            peers = [i for i in range(local_world) if nodes[i] == node]

        def core_of(c):     # hyper-thread siblings next to each other, so that a slice is made of whole cores
            try:
                return min(_cpulist(open("/sys/devices/system/cpu/cpu%d/topology/thread_siblings_list" % c).read()))
            except OSError:
                return c
        cpus.sort(key=lambda c: (core_of(c), c))
        me = peers.index(local_rank)
        per = len(cpus) // len(peers)
        if per < 2:
            return None
        mine = cpus[me * per:(me + 1) * per]
        os.sched_setaffinity(0, mine)
        os.environ["VM_HOST_THREADS"] = str(max(2, min(int(os.environ.get("VM_HOST_THREADS", "64")), len(mine))))
        return "rank bound to %d cores of NUMA node %d (its GPU's)" % (len(mine), node)
    except Exception as e:      # no sysfs / nvml in this container: run unpinned
        sys.stderr.write("pinning skipped: %s\n" % e)
        return None


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    sample = min(args.reads, args.cpu_sample)
    ref, reads, _note = cpu_inputs(sample)
    arm = CpuArm(ref, cores)
    rate = None
    for _ in range(max(args.warmup, 1)):
        wn = min(len(reads), max(cores * 8, 64))
        _, wdt = arm.run(reads[:wn])
        rate = wn / max(wdt, 1e-6)                 # reads per second of the warmed pool
    # a bounded sample per step: the K timed steps together stay within ~75 s whatever K the caller asks for
    budget_s = float(os.environ.get("VM_REF_BUDGET_S", "75"))
    sample = max(min(sample, 256), min(sample, int(rate * budget_s / max(args.steps, 1))))
    reads = reads[:sample]
    vals, t_all = [], 0.0
    for _ in range(args.steps):
        v, dt = arm.run(reads)
        vals.append(v)
        t_all += dt
    arm.close()
    v = float(np.mean(vals))
    emit({
        "impl": "reference", "metric": "aligned_gbp_per_s", "value": v, "unit": "Gbp/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * t_all / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+i32", "data": "synthetic",
        "config": {"workload": WL["name"], "reads_per_step": sample, "read_len": WL["read_len"], "err": WL["err"], "ref_len": WL["ref_len"]},
        "cpu_baseline": {"value": v, "unit": "Gbp/s", "cores": cores, "kind": "port",
                         "sample": "%d reads of the workload per step; oracle port (C stages + Python glue), %d persistent worker "
                                   "processes, warmed" % (sample, cores)},
        "e2e": {"value": v, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def make_asm_contig(ref, rank, length):
    """One contig per rank: a slice of a reference contig with a 20 kb inversion, a 5 kb deletion, a 3 kb insertion and
    0.1 % divergence."""
    import synth
    rng = np.random.default_rng(WL["read_seed"] + rank)
    src = np.frombuffer(ref[rank % len(ref)][1].encode(), dtype=np.uint8)
    st = int(rng.integers(0, len(src) - length - 10000))
    src = src[st:st + length + 5000]
    a, b, c = length // 5, int(length * 0.45), int(length * 0.7)
    parts = [src[:a], synth._COMP[src[a:a + 20000]][::-1], src[a + 20000:b], src[b + 5000:c], synth.random_seq(rng, 3000), src[c:]]
    return synth.mutate(rng, np.concatenate(parts), WL["err"], WL["ratio"]).tobytes().decode()


def run_asm(args):
    """--workload cfg4: contigs through `-mode asm` (vacmap_b200.asm.assembly_align: Python host loop like the reference's,
    every hot loop a CUDA entry point).  A step = one contig per GPU, host sequence in, record rows out."""
    import torch
    import vacmap_b200 as vb
    import vacmap_b200.__main__ as cli
    from vacmap_b200 import asm, shard
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    args.warmup = max(args.warmup, 3)
    ref = make_workload(0, ref_only=True)
    contig = make_asm_contig(ref, rank, WL["read_len"])
    ctx = vb._lib.Context(local_rank)
    t_ix = time.perf_counter()
    if dist is not None:
        ix = shard.broadcast_index(vb.Index(ref, w=WL["w"], k=WL["k"], ctx=ctx) if rank == 0 else None, ctx=ctx, device=local_rank)
    else:
        ix = vb.Index(ref, w=WL["w"], k=WL["k"], ctx=ctx)
    index_s = time.perf_counter() - t_ix
    opt = cli.options_from(cli.build_parser().parse_args(["-ref", "x", "-read", "x", "-mode", "asm", "--H", "--fakecigar"]))
    rows = None
    for _ in range(args.warmup):
        rows = asm.assembly_align("ctg%d" % rank, contig, ix, opt, ctx=ctx)
    sampler = ClockSampler(list(range(int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    asm.STATS.clear()
    l0 = ctx.kernel_launches
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rows = asm.assembly_align("ctg%d" % rank, contig, ix, opt, ctx=ctx)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    launches = ctx.kernel_launches - l0
    aligned = len(contig) * args.steps if rows else 0
    t = torch.tensor([wall], dtype=torch.float64, device="cuda")
    tot = torch.tensor([aligned, len(rows)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank != 0:
        return
    wall_max = float(t.cpu()[0])
    value = float(tot.cpu()[0]) / wall_max / 1e9
    stages = {k: round(1000 * v / args.steps, 1) for k, v in sorted(asm.STATS.items())}
    stages["python_host_loop"] = round(1000 * wall / args.steps - sum(stages.values()), 1)
    cpu = None
    if not args.no_cpu:
        # the oracle's restatement of the contig path on one core, on a 600 kb contig of a 5 Mb reference (the oracle's
        # index build is single-threaded)
        import oracle
        import oracle.asm as oasm
        import synth
        small_ref = synth.make_reference(WL["ref_seed"], 5_000_000)
        keep = WL["read_len"]
        WL["read_len"] = 600_000
        small = make_asm_contig(small_ref, 0, 600_000)
        WL["read_len"] = keep
        import oracle.pipeline as opl
        ox = oracle.Index(small_ref, w=WL["w"], k=WL["k"])
        octg = opl.Contigs([n for n, _ in small_ref], [s_ for _, s_ in small_ref])
        tc = time.perf_counter()
        oasm.assembly_align("ctg", small, ox, octg, opt)
        dt = time.perf_counter() - tc
        cpu = {"value": len(small) / dt / 1e9, "unit": "Gbp/s", "cores": 1, "kind": "port",
               "sample": "one 600 kb contig vs a 5 Mb reference, oracle port of the contig path (C stages + Python glue), one core, %.1f s" % dt}
    peak, peak_src = peaks()
    emit({"metric": "aligned_gbp_per_s", "value": value, "unit": "Gbp/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
          "ms_per_step": 1000 * wall_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+i32",
          "data": "synthetic",
          "config": {"workload": WL["name"], "contig_len": len(contig), "contigs_per_gpu_per_step": 1, "ref_len": WL["ref_len"], "mode": "asm",
                     "k": WL["k"], "w": WL["w"], "index_build_s": round(index_s, 2), "records_per_contig": len(rows),
                     "l2": "a contig's anchors, hits and direction matrices exceed the 126 MB L2",
                     "note": "no resident variant: a contig is processed batch by batch by the Python host loop, so value == e2e"},
          "e2e": {"value": value, "unit": "Gbp/s", "h2d_bytes_per_step": len(contig), "d2h_bytes_per_step": int(sum(len(r[8]) for r in rows))},
          "gpu_launches": int(launches / max(args.steps, 1)), "stage_ms_per_step": stages,
          "roofline": {"bound": "hbm", "kernel": None, "achieved": None, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": None,
                       "traffic": None,
                       "note": "host-bound: the time outside the CUDA entry points (python_host_loop) dominates; the kernels are the per-read "
                               "path's (roofline in the default bench line)"},
          "cpu_baseline": cpu, "clocks": clocks})


KERNEL_BYTES_NOTE = {
    "k_fill": "B = sum(q+t) sequence bytes + 4 B per CIGAR op (SURVEY 8d's compulsory terms); the direction bytes of the band cells "
              "(one 128-byte line per step and direction word, written once, re-read along the path) exceed shared memory and "
              "spill to HBM: reported apart as direction_spill_bytes / frac_with_direction_spill",
    "k_ed_upper": "B = sum(q+t) sequence bytes (every base is read once, by the gap DP or by the anchor check)",
    "k_seed": "B = read bytes + 16 B per anchor out",
    "k_edit_distance": "B = sum(q+t) sequence bytes (bit-vector state stays in registers/SMEM)",
    "k_reseed_hits": "B = read bytes x2 strands + 8 B per hit written (x2 launches: count + fill)",
    "k_reseed_merge": "B = 8 B per hit read + 16 B per anchor out",
    "chain_local_kernels": "B = 28 B per anchor (16 in + 8 S + 4 P)",
    "chain_global_kernels": "B = 28 B per anchor (16 in + 8 S + 4 P)",
    "k_extend": "B = sum(q+t) sequence bytes",
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU per step (0 = the workload's: configs[1] 10k)")
    ap.add_argument("--workload", default="cfg1", choices=sorted(WORKLOADS), help="cfg1 = BASELINE configs[1] (the headline)")
    ap.add_argument("--cpu-sample", type=int, default=2048, help="reads in the bounded CPU baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workers", type=int, default=0, help="sub-batches in flight per GPU (0 = library default)")
    ap.add_argument("--chunk", type=int, default=0, help="reads per sub-batch (0 = automatic)")
    ap.add_argument("--ahead", type=int, default=0, help="steps submitted ahead of the one being collected (0 = the workload's: 3)")
    args = ap.parse_args()
    global WL
    WL = WORKLOADS[args.workload]
    if args.reads <= 0:
        args.reads = WL["reads"]
    if args.ahead <= 0:
        args.ahead = WL.get("ahead", 3)
    if args.workers <= 0:
        args.workers = WL.get("workers", 0)
    # stdout carries exactly one line, the JSON: everything else that writes to fd 1 while the run lasts (NCCL's
    # version banner, library chatter) is sent to stderr, and the line goes out through the saved descriptor
    global _STDOUT
    sys.stdout.flush()
    _STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.workload == "cfg4" and args.impl != "reference":
        run_asm(args)
        return
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    # one process per GPU: every rank gets its share of the host cores for the glue (pool size is read at load time),
    # and those cores are the ones of the NUMA node its GPU hangs off (staging copies stay on the near socket)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    pinning = None
    if local_world > 1:
        os.environ.setdefault("VM_HOST_THREADS", str(max(2, (os.cpu_count() or 1) // local_world)))
        if os.environ.get("VM_BENCH_PIN", "1") != "0":
            pinning = pin_to_gpu_node(int(os.environ.get("LOCAL_RANK", "0")), local_world)
    os.environ.setdefault("NCCL_DEBUG", "WARN")       # NCCL's version banner goes to stdout; this script prints one JSON line there
    import torch
    import vacmap_b200 as vb

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # reads shard across ranks (no data-path collective): weak scaling, fixed work per GPU
    ref, reads, cat, off = make_workload(args.reads, rank)
    if dist is not None:
        from vacmap_b200 import shard
    ctx = vb._lib.Context(local_rank)
    t_ix = time.perf_counter()
    if dist is not None and os.environ.get("VM_BROADCAST_INDEX", "1") != "0":
        # the index is built once, on rank 0's GPU, and the BUILT tables go out over NCCL (HBM to HBM)
        ix = shard.broadcast_index(vb.Index(ref, w=WL["w"], k=WL["k"], ctx=ctx) if rank == 0 else None, ctx=ctx, device=local_rank)
    else:
        ix = vb.Index(ref, w=WL["w"], k=WL["k"], ctx=ctx)
    index_s = time.perf_counter() - t_ix
    al = vb.Aligner(ix, vb.default_option(WL["mode"]), WL["mode"], workers=args.workers, chunk_reads=args.chunk)
    bases = int(off[-1])

    # ---- device-resident: reads already in HBM when the timed region starts ----
    al.upload_reads(cat, off)
    from collections import deque
    # warm-up with the very submission pattern of the timed loop (jobs in flight, every worker's arenas grown)
    warm = deque()
    for _ in range(args.warmup):
        warm.append(al.submit_packed(cat, off, resident=True))
        if len(warm) > args.ahead:
            al.wait(warm.popleft())
    while warm:
        rec_off, recs, cig = al.wait(warm.popleft())
    sampler = ClockSampler(list(range(local_world))) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    l0 = ctx.kernel_launches
    stage = {}
    aligned = 0
    t0 = time.perf_counter()
    cpu0 = time.process_time()
    # steps are submitted ahead of their collection (vm_align_submit / vm_align_wait, `--ahead` jobs in flight): while
    # step s drains, the next ones are already seeding, as a stream of super-batches would; every step's work
    # completes inside the timed region
    inflight = deque()

    def collect():
        nonlocal aligned, rec_off, recs, cig
        rec_off, recs, cig = al.wait(inflight.popleft())
        for k, v in al.last_stage_ms.items():
            stage[k] = stage.get(k, 0.0) + v
        aligned += int(np.diff(off)[np.diff(rec_off) > 0].sum())

    rec_off = recs = cig = None
    for _ in range(args.steps):
        inflight.append(al.submit_packed(cat, off, resident=True))
        if len(inflight) > args.ahead:
            collect()
    while inflight:
        collect()
    barrier()
    wall = time.perf_counter() - t0
    cpu_busy = (time.process_time() - cpu0) / max(wall, 1e-9)      # host cores kept busy by this rank
    clocks = sampler.stop() if sampler else None
    launches = ctx.kernel_launches - l0

    # ---- end to end: host reads in, host records out, every step ----
    # the step's reads sit in page-locked host memory (vm_host_alloc; where a read parser would put them): their
    # upload is a DMA on the sub-batch's own stream, inside the timed region
    cat_pin = ctx.pinned_bytes(len(cat))
    cat_pin[:] = np.frombuffer(cat, dtype=np.uint8)
    for _ in range(2):
        warm.append(al.submit_packed(cat_pin, off))
    while warm:
        al.wait(warm.popleft())
    barrier()
    t0 = time.perf_counter()
    aligned_e2e = 0
    for _ in range(args.steps):
        inflight.append(al.submit_packed(cat_pin, off))      # host reads in ...
        if len(inflight) > args.ahead:
            rec_off, recs, cig = al.wait(inflight.popleft())     # ... host records out
            aligned_e2e += int(np.diff(off)[np.diff(rec_off) > 0].sum())
    while inflight:
        rec_off, recs, cig = al.wait(inflight.popleft())
        aligned_e2e += int(np.diff(off)[np.diff(rec_off) > 0].sum())
    barrier()
    e2e_wall = time.perf_counter() - t0
    d2h = recs.nbytes + cig.nbytes + rec_off.nbytes
    free_b, total_b = torch.cuda.mem_get_info()
    hbm_used_gb = round((total_b - free_b) / 1e9, 1)      # index + every worker's arenas, after both timed legs

    t = torch.tensor([wall, e2e_wall], dtype=torch.float64, device="cuda")
    tot = torch.tensor([aligned, aligned_e2e, len(recs)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        # Records stay on the rank that made them (per-rank SAM emit, SURVEY 8e: the reference's output order is
        # unspecified for more than one worker); the optional single-stream gather is timed apart, warmed, in its
        # steady state (shard.gather_records: sizes by all_gather, unpadded send / recv)
        shard.gather_records(rec_off, recs, cig)
        tg = time.perf_counter()
        gathered = shard.gather_records(rec_off, recs, cig)
        gather_ms = 1000 * (time.perf_counter() - tg)
        if rank == 0:
            assert len(gathered[1]) == int(tot[2].item())
    else:
        gather_ms = 0.0
    # ---- SAM text of one step's records on the host threads (vm_sam_batch; no device work): reported apart, the emitter
    # is downstream of the metric ----
    sam_text = None
    if rank == 0:
        try:
            from vacmap_b200 import sam as _sam
            table = _sam.ContigTable(ix)
            sam_opt = {"H": False, "fakecigar": False, "rg-id": "1"}
            batch = [(n, None) for n, _ in reads]
            _sam.batch_text(batch, rec_off, recs, cig, table, sam_opt, packed_seqs=(cat, off))
            ts = time.perf_counter()
            text, _ = _sam.batch_text(batch, rec_off, recs, cig, table, sam_opt, packed_seqs=(cat, off))
            dts = time.perf_counter() - ts
            sam_text = {"ms_per_step": round(1000 * dts, 1), "gbp_per_s": round(bases / dts / 1e9, 3), "text_bytes": len(text),
                        "threads": int(os.environ.get("VM_HOST_THREADS", os.cpu_count() or 1)),
                        "note": "SAM lines of one step's records (NM, SA, RG; soft-clipped supplementary records carry the whole read) "
                                "written by the library's host threads, byte-identical to the reference's emitter; not part of `value` / `e2e`"}
            del text
        except Exception as e:      # never let the side measurement break the line
            sam_text = {"error": str(e)[:200]}
    # ---- roofline leg: one lock-step pass (one worker, one stream), so every kernel is timed alone by the CUDA
    # events the library records on its launching stream; the pipelined legs above overlap kernels of several
    # workers, which stretches their individual durations ----
    solo = {}
    solo_note = None
    if rank == 0:
        al1 = vb.Aligner(ix, vb.default_option(WL["mode"]), WL["mode"], workers=1)
        try:
            for _ in range(2):
                al1.align_packed(cat, off, resident=True)
                solo = dict(al1.last_stage_ms)
        except Exception as e:      # e.g. not enough HBM left for whole-batch arenas beside the workers' (GRCh38-sized index)
            sys.stderr.write("lock-step roofline leg failed (%s): kernel times taken from the pipelined steps\n" % e)
            solo = {k: v / args.steps for k, v in stage.items()}
            solo_note = "kernel times from the pipelined steps (stretched by the overlap of the workers): the lock-step pass did not fit"

    wall_max, e2e_max = [float(x) for x in t.cpu()]
    aligned_all, aligned_e2e_all, nrec_all = [float(x) for x in tot.cpu()]

    if rank == 0:
        value = aligned_all / wall_max / 1e9
        e2e = aligned_e2e_all / e2e_max / 1e9
        peak, peak_src = peaks()
        per_step = {k: v / args.steps for k, v in stage.items()}
        kern = {k: v for k, v in solo.items() if k.startswith("k_") or k.endswith("_kernels")}
        top = max(kern, key=kern.get) if kern else None
        counts = {k: per_step.get(k, 0.0) for k in ("n_fill_cells", "n_fill_bases", "n_fill_jobs", "n_fill_band_jobs", "n_fill_band_redo", "n_fill_dir_bytes",
                                                      "n_ed_cells", "n_ed_upper_jobs", "n_reseed_hits", "n_chain_anchors", "n_chain_opcount", "n_syncs")}
        n_ops = float(len(cig))
        # ALGORITHMIC bytes per step: what the stage must move whatever the implementation
        alg_bytes = {"k_fill": counts["n_fill_bases"] + 4.0 * n_ops,
                     "k_edit_distance": 2.0 * bases, "k_ed_upper": 2.0 * bases,
                     "k_reseed_hits": 2.0 * bases + 8.0 * counts["n_reseed_hits"],
                     "k_reseed_merge": 8.0 * counts["n_reseed_hits"] + 16.0 * counts["n_reseed_hits"] / 4,
                     "chain_local_kernels": 28.0 * counts["n_chain_anchors"], "chain_global_kernels": 28.0 * counts["n_chain_anchors"],
                     "k_seed": 1.0 * bases + 16.0 * counts["n_chain_anchors"], "k_extend": 0.0}
        roof = None
        traffic_file = None
        for name in ("r2_kernel_traffic.json", "r1_kernel_traffic.json"):
            if os.path.exists(os.path.join(ROOT, "profiles", name)):
                traffic_file = os.path.join(ROOT, "profiles", name)
                break

        def traffic_of(kernel):
            if not traffic_file:
                return None
            tj = json.load(open(traffic_file)).get(kernel)
            if not tj:      # DRAM bytes per unit of work from the committed `ncu --set full` capture, scaled to this launch
                return None
            return tj["dram_bytes_per_unit"] * counts.get(tj["unit_count"], 0.0)
        if top:
            secs = kern[top] / 1000.0
            achieved = alg_bytes.get(top, 0.0) / secs / 1e9 if secs > 0 else 0.0
            spill = counts["n_fill_dir_bytes"] if top == "k_fill" else 0.0
            roof = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic_of(top),
                    "traffic_source": ("profiles/" + os.path.basename(traffic_file)) if traffic_file and traffic_of(top) is not None else None,
                    "kernel_ms": kern[top],
                    "kernel_ms_in_pipeline_per_step": per_step.get(top),
                    "timing": solo_note or "CUDA events on the launching stream, lock-step pass (one worker) after the timed region; one "
                                           "'launch' = the kernel's launches of one step (one per capacity class)",
                    "bytes": alg_bytes.get(top, 0.0), "bytes_formula": KERNEL_BYTES_NOTE.get(top, ""),
                    "direction_spill_bytes": spill,
                    "frac_with_direction_spill": ((alg_bytes.get(top, 0.0) + spill) / secs / 1e9 / peak) if secs > 0 else None,
                    "gcups_full_matrix_equivalent": (counts["n_fill_cells"] / secs / 1e9) if top == "k_fill" and secs > 0 else None,
                    "all_kernels_ms": {k: round(v, 3) for k, v in sorted(kern.items())},
                    # wall time of every stage of the lock-step pass (device-side glue kernels are inside front_device / extend_device)
                    "lockstep_stage_ms": {k: round(v, 3) for k, v in sorted(solo.items())
                                          if not k.startswith(("n_", "c_", "k_", "cpu_", "t_")) and not k.endswith("_kernels") and v >= 0.05},
                    "note": "integer DP wavefront: issue / ALU-pipe bound (ncu, 48-row class: issue 71 %, ALU 67 %, DRAM 15 %), not HBM bound "
                            "(SURVEY 8d); the HBM fraction is reported because north_star asks for it"}
        # the chaining kernels (the ones north_star names): both byte counts of SURVEY 8d
        chain_ms = kern.get("chain_global_kernels", 0.0) + kern.get("chain_local_kernels", 0.0)
        roof_chain = None
        if chain_ms > 0:
            b_chain = 28.0 * counts["n_chain_anchors"]
            b_alg = 28.0 * counts["n_chain_opcount"] + b_chain
            roof_chain = {"bound": "hbm", "kernel": "chain_global_kernels + chain_local_kernels (argsort replay + exact / fast DP)",
                          "kernel_ms": chain_ms, "peak": peak, "unit": "GB/s",
                          "B_chain": b_chain, "achieved": b_chain / (chain_ms / 1e3) / 1e9, "frac": b_chain / (chain_ms / 1e3) / 1e9 / peak,
                          "B_chain_alg": b_alg, "achieved_alg": b_alg / (chain_ms / 1e3) / 1e9,
                          "frac_alg": b_alg / (chain_ms / 1e3) / 1e9 / peak,
                          "anchors": counts["n_chain_anchors"], "opcount": counts["n_chain_opcount"],
                          "traffic": traffic_of("chain"),
                          "note": "B_chain = 28 B per anchor (16 in + 8 S + 4 P), the compulsory bytes; B_chain_alg = opcount x 28 + 28 n, what the "
                                  "reference's predecessor loop moves unstaged (opcount = its own counter, exact here).  The recurrence is serial "
                                  "along the anchors of a read: latency-bound, a launch lasts as long as its slowest read (ncu)"}
        line = {"metric": "aligned_gbp_per_s", "value": value, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1000 * wall_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64+i32", "data": "synthetic",
                "config": {"workload": WL["name"], "reads_per_gpu_per_step": args.reads, "read_len": WL["read_len"], "err": WL["err"],
                           "ref_len": WL["ref_len"], "mode": WL["mode"], "k": WL["k"], "w": WL["w"], "index_build_s": round(index_s, 2), "pinning": pinning, "hbm_used_gb": hbm_used_gb,
                           "l2": "per-step working set (reads 2x%.0f MB + anchors, hits, direction matrices >1 GB) exceeds "
                                 "the 126 MB L2" % (bases / 1e6),
                           "sharding": "reads split across ranks (no data-path collective); index built on rank 0's GPU and broadcast as built tables over NCCL; records "
                                       "emitted per rank (SURVEY 8e) -- gather_ms = one warmed gather of a step's records to rank 0, optional",
                           "pipelining": "steps submitted %d ahead of their collection (vm_align_submit / vm_align_wait), all K "
                                         "steps complete inside the timed region" % args.ahead},
                "e2e": {"value": e2e, "unit": "Gbp/s", "h2d_bytes_per_step": len(cat) + off.nbytes, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches), "host_cores_busy": round(cpu_busy, 2), "host_cores": os.cpu_count(),
                "records_per_step": nrec_all, "gather_ms": round(gather_ms, 2), "sam_text": sam_text,
                "stage_ms_per_step": {k: round(v, 3) for k, v in per_step.items() if not k.startswith("n_")},
                "work_per_step": counts,
                "roofline": roof, "roofline_chain": roof_chain, "clocks": clocks}
        line["cpu_baseline"] = None
        if not args.no_cpu and world == 1:      # the CPU leg belongs to the N = 1 line only (the other ranks' boxes would idle through it)
            cores = os.cpu_count() or 1
            sample = min(args.reads, args.cpu_sample)
            ref_c, reads_c, note = cpu_inputs(sample, ref)
            arm = CpuArm(ref_c, cores)
            arm.run(reads_c[:max(cores * 8, 64)])
            v, dt = arm.run(reads_c[:sample])
            arm.close()
            line["cpu_baseline"] = {"value": v, "unit": "Gbp/s", "cores": cores, "kind": "port",
                                    "sample": "%d reads of the workload, oracle port (C stages + Python glue), %d persistent worker "
                                              "processes (warmed), %.1f s" % (sample, cores, dt)}
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
