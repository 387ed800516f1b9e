#!/usr/bin/env python
"""Experiment builds of the engine next to the product library:
    python tools/build_variant.py NAME [--rev GIT_REV] [--flags "-DX=1 ..."]
-> baryonyx_b200/exp/lib_NAME.so (git-ignored; select it with BXG_LIB=... on the GPU box).
--rev builds the csrc/ tree of another commit (same ABI assumed)."""
import argparse, os, shutil, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baryonyx_b200 import build as B

ap = argparse.ArgumentParser()
ap.add_argument("name")
ap.add_argument("--rev", default=None)
ap.add_argument("--flags", default="")
a = ap.parse_args()
work = tempfile.mkdtemp(prefix="bxg_variant_")
csrc = os.path.join(work, "baryonyx_b200", "csrc")
os.makedirs(csrc)
os.makedirs(os.path.join(work, "include"))
files = [os.path.join("baryonyx_b200", "csrc", f) for f in os.listdir(B.CSRC)] + [os.path.join("include", "bxg.h")]
for f in files:
    dst = os.path.join(work, f)
    if a.rev:
        try:
            data = subprocess.check_output(["git", "show", f"{a.rev}:{f}"], cwd=ROOT)
        except subprocess.CalledProcessError:
            continue
        open(dst, "wb").write(data)
    else:
        shutil.copy(os.path.join(ROOT, f), dst)
ccbin = ["-ccbin", "/usr/bin/g++"]
procs, objs = [], []
for src, obj, defs in B.UNITS:
    o = os.path.join(work, obj)
    objs.append(o)
    cmd = [B.nvcc()] + B.NVCC_FLAGS + defs + a.flags.split() + ["-c", os.path.join(csrc, src), "-o", o] + ccbin
    procs.append(subprocess.Popen(cmd))
for p in procs:
    if p.wait():
        sys.exit("compile failed")
out = os.path.join(ROOT, "baryonyx_b200", "exp", f"lib_{a.name}.so")
subprocess.check_call([B.nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs + ccbin)
shutil.rmtree(work)
print(out)
