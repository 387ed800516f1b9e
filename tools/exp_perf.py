#!/usr/bin/env python
"""Experiment driver (GPU box): one bench step of the SPP workload with the
launch knobs taken from the environment; prints time and the per-path
counters.  usage: python tools/exp_perf.py [--lanes N] [--limit K] [--workload W]
The cycle counters need a profiling build: BXG_NVCC_FLAGS=-DBXG_TIMERS=1 python -m baryonyx_b200.build --force"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from baryonyx_b200 import capi

ap = argparse.ArgumentParser()
ap.add_argument("--lanes", type=int, default=1024)
ap.add_argument("--limit", type=int, default=100)
ap.add_argument("--workload", default="spp")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
inst, name = bench.make_instance(a.workload)
p = capi.default_params(); p.limit = a.limit
cp = capi.clamp_params(p)
pb = capi.Problem(inst, device=0)
bt = pb.batch(a.lanes, cp)
best = None
for r in range(a.reps):
    bt.reset_counters()
    seeds = (np.arange(a.lanes, dtype=np.uint64) + 1 + r * a.lanes).astype(np.uint32)
    bt.begin(seeds, reset_solver=True)
    bt.solve()
    ms = bt.last_ms()
    c = bt.counters(); pc = bt.path_counters()
    rec = {"ms": ms, "updates": c["updates"], "Mupd_s": c["updates"] / ms / 1e3,
           "alg_GBs": c["algorithmic_bytes"] / ms / 1e6, "paths": pc["updates"],
           "cyc_per_upd": [round(cy / max(1, u)) for u, cy in zip(pc["updates"], pc["cycles"])],
           "phase_cyc": [round(cy / max(1, pc["updates"][2] + pc["updates"][3])) for cy in pc["phase_cycles"]]}
    if best is None or rec["ms"] < best["ms"]:
        best = rec
best.update({"lanes": a.lanes, "env": {k: v for k, v in os.environ.items() if k.startswith("BXG_")}})
print(json.dumps(best), flush=True)
