import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from baryonyx_b200 import capi, synth
for name, kw in (("default", {}), ("linear only", dict(ex_frac=0.0, bb_frac=0.0)), ("exhaustive only", dict(ex_frac=0.10, bb_frac=0.0)), ("bb only", dict(ex_frac=0.0, bb_frac=0.10))):
    inst = synth.telebus(**kw)
    p = capi.default_params(); p.limit = 50; p.order = 3
    cp = capi.clamp_params(p)
    pb = capi.Problem(inst, device=0)
    bt = pb.batch(1024, cp)
    for rep in range(2):
        bt.reset_counters()
        bt.begin((np.arange(1024) + 1).astype(np.uint32), reset_solver=True)
        bt.solve()
        ms = bt.last_ms(); c = bt.counters()
    print(f"{name:16s} {ms:9.1f} ms  {c['updates']/ms/1e3:8.2f} M updates/s  updates {c['updates']}  sweeps {c['sweeps']}", flush=True)
    bt.close(); pb.close()
