"""Builds the native library in-tree: baryonyx_b200/libbaryonyx_b200.so.

nvcc cross-compiles for sm_100a without a GPU.  The .so is git-ignored but
travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbaryonyx_b200.so")
# bxg_kernels.cu is compiled once per solver family (-DBXG_FAM=0|1|2), the
# translation units in parallel
UNITS = [("bxg_api.cu", "bxg_api.o", []),
         ("bxg_kernels.cu", "bxg_kernels_f0.o", ["-DBXG_FAM=0"]),
         ("bxg_kernels.cu", "bxg_kernels_f1.o", ["-DBXG_FAM=1"]),
         ("bxg_kernels.cu", "bxg_kernels_f2.o", ["-DBXG_FAM=2"])]
SOURCES = ["bxg_api.cu", "bxg_kernels.cu"]
OBJDIR = os.path.join(HERE, "build")
HEADERS = ["bxg_device.cuh", "bxg_kernels.cuh", "bxg_zsolver.cuh", "bxg_population.cuh", "bxg_lp.hpp", "bxg_preprocess.hpp", os.path.join("..", "..", "include", "bxg.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # deterministic fp64 mode: never contract a*b+c into an FMA
    "-fmad=false",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-O2",
    # ptxas works on the kernels of a translation unit in parallel (each unit
    # holds ~20 instantiations of the sweep kernels): the SASS is byte for byte
    # the one a single-threaded ptxas writes (checked with cuobjdump -sass),
    # the Z-family unit takes 3 minutes instead of 9
    "-Xptxas", "--split-compile=0",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the engine has no CPU build")
    return exe


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [__file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    # the image exports CXX=/opt/gcc/bin/g++ (static libstdc++); nvcc must use
    # the system host compiler
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    procs = []
    for src, obj, defs in UNITS:
        # BXG_NVCC_FLAGS: extra flags for profiling builds, e.g. -DBXG_TIMERS=1
        # (per-path / per-phase cycle counters, tools/exp_perf.py)
        extra = os.environ.get("BXG_NVCC_FLAGS", "").split()
        cmd = ([nvcc()] + NVCC_FLAGS + defs + extra + (["-Xptxas", "-v"] if verbose else []) +
               ["-c", os.path.join(CSRC, src), "-o", os.path.join(OBJDIR, obj)] + ccbin)
        log = open(os.path.join(OBJDIR, obj + ".log"), "w")
        procs.append((cmd, log, subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT)))
    failed = None
    for cmd, log, p in procs:
        rc = p.wait()
        log.close()
        if verbose or rc != 0:
            with open(log.name) as f:
                print(f.read())
        if rc != 0 and failed is None:
            failed = subprocess.CalledProcessError(rc, cmd)
    if failed is not None:
        raise failed
    subprocess.check_call([nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] +
                          [os.path.join(OBJDIR, obj) for _, obj, _ in UNITS] + ccbin)
    return LIB


if __name__ == "__main__":
    import sys

    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
