"""Builds libmotb200.so (the product library) in-tree with nvcc for sm_100a.

    python -m motcpp_b200.build            # build if sources are newer than the .so
    python -m motcpp_b200.build --force

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmotb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                      # exact-arithmetic contract: no implicit FMA contraction anywhere
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]
OBJ_DIR = os.path.join(HERE, "build")   # git-ignored and gpurun-ignored: only the linked .so travels


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libmotb200.so cannot be built (there is no CPU fallback)")
    return exe


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                  glob.glob(os.path.join(ROOT, "include", "*.h")))


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    # one translation unit per fused engine kernel family + the C ABI with the stand-alone kernels, compiled in parallel
    units = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_time = max(os.path.getmtime(h) for h in sources() if not h.endswith(".cu"))

    def compile_unit(cu):
        obj = os.path.join(OBJ_DIR, os.path.basename(cu)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(hdr_time, os.path.getmtime(cu)):
            return obj, "", 0
        cmd = [nvcc(), *NVCC_FLAGS, "-c", "-o", obj, cu]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        return obj, " ".join(cmd) + "\n" + proc.stdout + proc.stderr, proc.returncode

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as pool:
        results = list(pool.map(compile_unit, units))
    log = "".join(r[1] for r in results)
    rc = max(r[2] for r in results)
    if rc == 0:
        cmd = [nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *[r[0] for r in results], "-lcudart"]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + proc.stdout + proc.stderr
        rc = proc.returncode
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(log)
    if rc != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libmotb200.so")
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
