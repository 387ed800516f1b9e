"""Bisect helper: full-size rail lockstep (2 sweeps) in whatever repo root is the cwd."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
from baryonyx_b200 import capi, synth
from oracle import pyoracle as po
from tests import util
from tests.test_gpu_parity import to_capi

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
inst = synth.rail(m=int(4284 * scale), n=int(1092610 * scale))
p = util.params(init_policy_random=0.0)
cp = po.clamp_params(p)
op = po.OracleProblem(inst)
runs = [op.run(cp, s) for s in (5, 6)]
pb = capi.Problem(inst)
bt = pb.batch(2, to_capi(cp))
bt.begin([5, 6])
for s in range(2):
    rem = [int(v) for v in bt.sweep(0.0)]
    ref = [r.sweep(0.0) for r in runs]
    st_ok = []
    for l, r in enumerate(runs):
        a, b = r.state(), bt.state(l)
        d = {k: bool(np.array_equal(a[k], b[k])) for k in ("x", "pi", "P")}
        d["rng"] = a["rng"] == b["rng"]
        st_ok.append(d)
    print("scale", scale, "sweep", s, "engine", rem, "oracle", ref, st_ok, flush=True)
    if rem != ref:
        a, b = runs[0].state(), bt.state(0)
        d = np.nonzero(a["pi"] != b["pi"])[0]
        rl = np.diff(inst["row_ptr"])
        print("rows with different pi:", len(d), d[:10], "their lengths", rl[d[:10]], "max_row", rl.max())
        break
