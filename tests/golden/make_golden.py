#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled here
(oracle/_ref/libbxref.so, built by oracle/Makefile from /root/reference).

Each file holds one instance (merged constraints + objective, as the reference's
own parser and make_merged_constraints produced them from the LP text this
script wrote), the solver parameters, the lane seeds, and -- per seed -- the
state of the reference's Solver after every sweep of a fixed kappa schedule:
x words, pi, P (row-major), random-engine state, remaining constraints, plus
the result of a whole solve (status, x, value).  The fixtures travel to the GPU
box, where /root/reference does not exist.

    python tests/golden/make_golden.py        # needs oracle/_ref/libbxref.so
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from baryonyx_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from tests import util  # noqa: E402

PFIELDS = [n for n, _ in po.Params._fields_]

CASES = [
    # name, generator, params, seeds, sweeps, pushes_every
    ("spp_eq01", lambda: synth.spp(m=40, n=300, lo=3, hi=6, seed=2, cost_hi=5),
     dict(limit=120, kappa_step=0.002, pushes_limit=3, pushing_iteration_limit=5), [11, 12, 13], 30, 7),
    ("scp_ineq01", lambda: synth.scp(m=40, n=200, rows_per_col=5, seed=3, cost_hi=3),
     dict(limit=100, kappa_step=0.005, init_policy_random=0.0, pushes_limit=2, pushing_iteration_limit=4),
     [21, 22], 20, 6),
    ("mixed_eq101", lambda: synth.small_mixed(m=16, n=40, seed=9, family="101", equality=True, cost_hi=5),
     dict(limit=150, kappa_step=0.01, w=3, pushes_limit=4, pushing_iteration_limit=6), [31, 32, 33], 40, 9),
    ("mixed_ineq101_max", lambda: synth.small_mixed(m=14, n=36, seed=11, family="101", maximize=True, cost_hi=4),
     dict(limit=150, kappa_step=0.01, w=3, order=po.ORDERS["random-sorting"], pushes_limit=4,
          pushing_iteration_limit=6), [41, 42], 40, 8),
    ("mixed_ineq01_infeas", lambda: synth.small_mixed(m=20, n=50, seed=6, family="01", cost_hi=3, density=0.2),
     dict(limit=100, kappa_step=0.01, w=2, order=po.ORDERS["infeasibility-decr"], pushes_limit=3,
          pushing_iteration_limit=5), [7, 8], 30, 5),
    ("mixed_ineqZ", lambda: synth.small_mixed(m=14, n=36, seed=5, family="Z", cost_hi=6),
     dict(limit=150, kappa_step=0.01, w=3, pushes_limit=3, pushing_iteration_limit=5), [51, 52], 30, 7),
    ("mixed_eqZ", lambda: synth.small_mixed(m=12, n=30, seed=8, family="Z", equality=True, cost_hi=6),
     dict(limit=150, kappa_step=0.01, w=3, pushes_limit=3, pushing_iteration_limit=5), [61, 62], 30, 7),
]


def trajectory(run, m, p, sweeps, pushes_every):
    """Same kappa schedule as tests/util.lockstep (solver_functor's)."""
    xs, pis, Ps, rngs, rems = [], [], [], [], []
    kappa = p.kappa_min
    st = run.state()
    xs.append(st["x"]), pis.append(st["pi"]), Ps.append(st["P"]), rngs.append(st["rng"]), rems.append(-1)
    for s in range(sweeps):
        push = 1 if (pushes_every and s % pushes_every == pushes_every - 1) else 0
        if push:
            rem = run.sweep(p.pushing_k_factor * kappa, 1, p.pushing_objective_amplifier)
        else:
            rem = run.sweep(kappa)
        st = run.state()
        xs.append(st["x"]), pis.append(st["pi"]), Ps.append(st["P"]), rngs.append(st["rng"]), rems.append(rem)
        if s > p.w:
            kappa += p.kappa_step * (rem / m) ** p.alpha
        kappa = min(kappa, 0.999)
    return (np.stack(xs), np.stack(pis), np.stack(Ps), np.asarray(rngs, np.uint32), np.asarray(rems, np.int32))


def main():
    if not po.ref_available():
        raise SystemExit("oracle/_ref/libbxref.so is missing: run `make -C oracle ref` first")
    for name, gen, pkw, seeds, sweeps, pushes_every in CASES:
        inst = gen()
        p = util.params(**pkw)
        lp = util.lp_of(inst)
        rp = po.RefProblem(lp, p)
        os.unlink(lp)
        cp = rp.params()  # after the reference's own clamping
        rinst = rp.instance()
        out = {
            "sweeps": np.int32(sweeps), "pushes_every": np.int32(pushes_every),
            "seeds": np.asarray(seeds, np.uint32),
            "params": np.asarray([float(getattr(cp, f)) for f in PFIELDS], np.float64),
        }
        for k in ("row_ptr", "col", "factor", "cmin", "cmax", "obj_var", "obj_factor"):
            out["inst_" + k] = np.asarray(rinst[k])
        out["inst_scalars"] = np.asarray([rinst["m"], rinst["n"], rinst["problem_type"], rinst["maximize"]],
                                         np.int64)
        out["inst_cost_constant"] = np.float64(rinst["cost_constant"])
        for l, seed in enumerate(seeds):
            run = rp.run(seed)
            x, pi, P, rng, rem = trajectory(run, rp.m, cp, sweeps, pushes_every)
            out[f"tail_{l}"] = np.float64(run.tail_value())
            run.close()
            out[f"x_{l}"], out[f"pi_{l}"], out[f"P_{l}"], out[f"rng_{l}"], out[f"rem_{l}"] = x, pi, P, rng, rem
            sol = rp.solve(seed)
            out[f"solve_{l}"] = np.asarray([sol["status"], sol["remaining"]], np.int64)
            out[f"solve_x_{l}"] = sol["x"]
            out[f"solve_value_{l}"] = np.float64(sol["value"])
        rp.close()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        h = hashlib.sha256(open(path, "rb").read()).hexdigest()[:12]
        print(f"{name}: m={rinst['m']} n={rinst['n']} type={rinst['problem_type']} seeds={seeds} "
              f"sweeps={sweeps} -> {os.path.getsize(path)} bytes sha256 {h}")


if __name__ == "__main__":
    main()
