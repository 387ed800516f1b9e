#!/usr/bin/env python
"""Raw path / phase cycle counters of one solve run (needs a -DBXG_TIMERS=1 build, BXG_LIB=...):
python tools/exp_phases.py --workload csplib022 --lanes 296 --limit 20"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from baryonyx_b200 import capi

ap = argparse.ArgumentParser()
ap.add_argument("--lanes", type=int, default=1024)
ap.add_argument("--limit", type=int, default=100)
ap.add_argument("--workload", default="spp")
a = ap.parse_args()
inst, name = bench.make_instance(a.workload)
p = capi.default_params(); p.limit = a.limit
pb = capi.Problem(inst, device=0)
bt = pb.batch(a.lanes, capi.clamp_params(p))
for r in range(2):
    bt.reset_counters()
    bt.begin((np.arange(a.lanes, dtype=np.uint64) + 1).astype(np.uint32), reset_solver=True)
    bt.solve()
c = bt.counters(); pc = bt.path_counters()
print(json.dumps({"workload": name, "ms": bt.last_ms(), "updates": c["updates"], "lanes": a.lanes,
                  "paths": pc["updates"], "path_cycles": pc["cycles"], "phase_cycles": pc["phase_cycles"],
                  "cycles_per_lane": [cy / a.lanes for cy in pc["phase_cycles"]]}), flush=True)
