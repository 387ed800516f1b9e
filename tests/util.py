"""Shared helpers for the parity tests (test infrastructure)."""
from __future__ import annotations

import os
import tempfile

import numpy as np

from baryonyx_b200 import synth
from oracle import pyoracle as po

REF_TEST_DIR = "/root/reference/lib/test"


def params(**kw) -> po.Params:
    p = po.default_params()
    for k, v in kw.items():
        if k == "order" and isinstance(v, str):
            v = po.ORDERS[v]
        setattr(p, k, v)
    return p


def lp_of(inst) -> str:
    fd, path = tempfile.mkstemp(suffix=".lp", prefix="bx_")
    os.close(fd)
    synth.write_lp(inst, path)
    return path


def states_equal(a, b, what=("x", "pi", "P", "rng", "order_len")):
    for k in what:
        if k in ("rng", "order_len"):
            if a[k] != b[k]:
                return False, k
        elif not np.array_equal(a[k], b[k], equal_nan=(k in ("pi", "P"))):
            # NaN == NaN here: instances without an objective get delta =
            # DBL_MAX/2 (compute_delta, itm-common.hpp:917-933) and overflow
            return False, k
    if "order_len" in what:
        n = a["order_len"]
        if not np.array_equal(a["order"][:n], b["order"][:n]):
            return False, "order"
    return True, None


def lockstep(run_a, run_b, m, p, sweeps, pushes_every=0, order_pairs=False):
    """Drive two step-level runs with the kappa schedule of solver_functor
    (lib/src/itm-solver-common.hpp:135-166) and compare all state after every
    sweep.  Returns the number of sweeps executed."""
    kappa = p.kappa_min
    for s in range(sweeps):
        push = 1 if (pushes_every and s % pushes_every == pushes_every - 1) else 0
        if push:
            ra = run_a.sweep(p.pushing_k_factor * kappa, 1, p.pushing_objective_amplifier)
            rb = run_b.sweep(p.pushing_k_factor * kappa, 1, p.pushing_objective_amplifier)
        else:
            ra = run_a.sweep(kappa)
            rb = run_b.sweep(kappa)
        assert ra == rb, f"remaining differs at sweep {s}: {ra} vs {rb}"
        sa, sb = run_a.state(), run_b.state()
        if order_pairs:
            sa["order_len"] *= 2
            sb["order_len"] *= 2
        ok, k = states_equal(sa, sb)
        assert ok, f"state '{k}' differs at sweep {s}"
        if s > p.w:
            kappa += p.kappa_step * (ra / m) ** p.alpha
        if kappa > p.kappa_max:
            return s + 1
    return sweeps
