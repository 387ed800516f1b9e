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
SOURCES = ["bxg_api.cu", "bxg_kernels.cu"]
HEADERS = ["bxg_device.cuh", "bxg_kernels.cuh", "bxg_zsolver.cuh", os.path.join("..", "..", "include", "bxg.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # deterministic fp64 mode: never contract a*b+c into an FMA
    "-fmad=false",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-O2",
    "-shared",
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
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
        "-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    env = dict(os.environ)
    # the image exports CXX=/opt/gcc/bin/g++ (static libstdc++); nvcc must use
    # the system host compiler
    subprocess.check_call(cmd + ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else cmd, env=env)
    return LIB


if __name__ == "__main__":
    import sys

    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
