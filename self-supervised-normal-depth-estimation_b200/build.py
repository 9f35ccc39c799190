"""Builds libsnde.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m snde_b200.build            # or: __graft_entry__.build()

-fmad=false: the arithmetic order of csrc/snde_math.cuh is part of the parity contract
(fused multiply-adds only where written as fmaf()); no --use_fast_math.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libsnde.so")
SOURCES = ["snde_kernels.cu"]
HEADERS = ["snde_math.cuh", "snde_tile.cuh", "snde_loss_setup.cuh"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_fresh() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(INCLUDE, "snde.h")]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if is_fresh() and not force:
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-fmad=false", "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
           "-Xptxas", "-v" if verbose else "-O3", "-I", INCLUDE, "-I", CSRC,
           "-DSNDE_BUILDING=1", "-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
