#!/usr/bin/env python
"""Where the end-to-end step of bench.py spends its time (GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from baryonyx_b200 import capi

inst, _ = bench.make_instance("spp")
p = capi.default_params(); p.limit = 100
cp = capi.clamp_params(p)
L = 1024
for rep in range(3):
    t = [time.perf_counter()]
    pb = capi.Problem(inst, device=0); t.append(time.perf_counter())
    bt = pb.batch(L, cp); t.append(time.perf_counter())
    bt.begin((np.arange(L) + 1 + rep * L).astype(np.uint32)); t.append(time.perf_counter())
    out = bt.solve(); t.append(time.perf_counter())
    ln, _, _ = bt.argbest(); x = bt.best(ln); c = bt.counters(); t.append(time.perf_counter())
    bt.close(); pb.close(); t.append(time.perf_counter())
    names = ["problem_create", "batch_create", "begin", "solve", "results", "destroy"]
    print({n: round(1e3 * (b - a), 2) for n, a, b in zip(names, t, t[1:])}, flush=True)
