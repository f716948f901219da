"""Build libvacmap_b200.so in-tree with nvcc for sm_100a (no torch involved)."""
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvacmap_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # float64 score sums must reproduce the reference's evaluation order: no FMA contraction
    "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
    "--shared", "-cudart", "shared",
]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: libvacmap_b200.so cannot be built (there is no CPU fallback)")
    return p


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


OBJ = os.path.join(HERE, "csrc", "_obj")


def headers():
    return (glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.hpp")) +
            glob.glob(os.path.join(HERE, "..", "include", "*.h")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in sources() + headers())


def build_library(force=False, verbose=False):
    """One object per .cu (compiled in parallel, only when stale), then one link."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ, exist_ok=True)
    nvcc = nvcc_path()
    flags = [f for f in NVCC_FLAGS if f not in ("--shared",)]
    hdr_t = max(os.path.getmtime(h) for h in headers())
    log = []

    def one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj
        cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, res.stdout, res.stderr))
        log.append(res.stdout + res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(one, sources()))
    res = subprocess.run([nvcc] + NVCC_FLAGS + ["-o", LIB] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    if verbose:
        print("".join(log))
    return LIB


if __name__ == "__main__":
    import sys
    print(build_library(force="-f" in sys.argv, verbose="-v" in sys.argv))
