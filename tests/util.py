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


# ---------------------------------------------------------------------------
# golden vectors (tests/golden/*.npz, written by tests/golden/make_golden.py
# from the compiled reference)
# ---------------------------------------------------------------------------
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith("optimizer_"))


def load_golden(name):
    """-> (instance dict, clamped Params, npz)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    m, n, ptype, maximize = [int(v) for v in z["inst_scalars"]]
    inst = {
        "m": m, "n": n, "problem_type": ptype, "maximize": maximize,
        "cost_constant": float(z["inst_cost_constant"]),
    }
    for k in ("row_ptr", "col", "factor", "cmin", "cmax", "obj_var", "obj_factor"):
        inst[k] = z["inst_" + k]
    p = po.Params()
    for (fname, ftype), v in zip(po.Params._fields_, z["params"]):
        setattr(p, fname, float(v) if "double" in ftype.__name__ else int(v))
    return inst, p, z


def golden_kappa_schedule(z, p, m, lane):
    """Yields (sweep index, push flag, kappa) following the recorded remaining
    counts, exactly as make_golden.trajectory drove the reference."""
    kappa = p.kappa_min
    pushes_every = int(z["pushes_every"])
    rem = z[f"rem_{lane}"]
    for s in range(int(z["sweeps"])):
        push = 1 if (pushes_every and s % pushes_every == pushes_every - 1) else 0
        yield s, push, kappa
        if s > p.w:
            kappa += p.kappa_step * (int(rem[s + 1]) / m) ** p.alpha
        kappa = min(kappa, 0.999)
