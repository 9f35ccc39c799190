"""Builds libsnde.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m snde_b200.build            # or: __graft_entry__.build()

Arithmetic order is part of the parity contract (csrc/snde_math.cuh): the reference-signature kernels
(snde_ops.cu) are built with -fmad=false; the fused kernel (snde_fused.cu) with -fmad=true, its exact-order
chains written out with fmaf()/__fmul_rn()/__fadd_rn().  No --use_fast_math anywhere.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.environ.get("SNDE_LIB", os.path.join(HERE, "libsnde.so"))   # SNDE_LIB / SNDE_NVCC_DEFS: tuning experiments only
# (source, fp-contraction flag): the fused kernel spells out its exact-order arithmetic, the reference-signature
# kernels follow the reference op by op
SOURCES = [("snde_fused.cu", "-fmad=true"), ("snde_ops.cu", "-fmad=false"), ("snde_dn.cu", "-fmad=true")]
HEADERS = ["snde_math.cuh", "snde_strip.cuh", "snde_loss_setup.cuh"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_fresh() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in [s for s, _ in SOURCES] + HEADERS] + [os.path.join(INCLUDE, "snde.h")]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if is_fresh() and not force:
        return LIB
    common = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden", "-Xptxas", "-v",
              "-I", INCLUDE, "-I", CSRC, "-DSNDE_BUILDING=1"] + os.environ.get("SNDE_NVCC_DEFS", "").split()
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src, fmad in SOURCES:
        obj = os.path.join(HERE, "build", os.path.basename(LIB).replace(".so", "_") + src.replace(".cu", ".o"))
        res = subprocess.run(common + [fmad, "-c", "-o", obj, os.path.join(CSRC, src)], capture_output=True, text=True)
        if verbose:
            print(res.stderr)
        # register / spill report of every kernel: tests/test_abi_cpu.py holds the hot instance to zero spills
        with open(obj.replace(".o", ".ptxas.log"), "w") as f:
            f.write(res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + res.stderr[-4000:])
        objs.append(obj)
    res = subprocess.run([nvcc_path(), "-shared", "-o", LIB] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
