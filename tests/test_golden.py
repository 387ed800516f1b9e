"""Golden vectors of the compiled reference (tests/golden/*.npz): the C oracle
must reproduce them on any box (CPU test), and so must the CUDA engine through
the C ABI (-m gpu).  Bit-exact: x words, pi, P, engine state, remaining
constraints after every sweep; status / x / value of the whole solve."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as po
from tests import util

NAMES = util.golden_names()


def test_fixtures_present():
    assert len(NAMES) >= 7


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference_trajectory(name):
    inst, p, z = util.load_golden(name)
    op = po.OracleProblem(inst)
    for lane, seed in enumerate(z["seeds"]):
        op.set_tail_value(float(z[f"tail_{lane}"]))
        run = op.run(p, int(seed))
        st = run.state()
        assert np.array_equal(st["x"], z[f"x_{lane}"][0])
        for s, push, kappa in util.golden_kappa_schedule(z, p, inst["m"], lane):
            if push:
                rem = run.sweep(p.pushing_k_factor * kappa, 1, p.pushing_objective_amplifier)
            else:
                rem = run.sweep(kappa)
            st = run.state()
            assert rem == int(z[f"rem_{lane}"][s + 1]), (name, lane, s)
            assert np.array_equal(st["x"], z[f"x_{lane}"][s + 1]), (name, lane, s)
            assert np.array_equal(st["pi"], z[f"pi_{lane}"][s + 1], equal_nan=True), (name, lane, s)
            assert np.array_equal(st["P"], z[f"P_{lane}"][s + 1], equal_nan=True), (name, lane, s)
            assert st["rng"] == int(z[f"rng_{lane}"][s + 1]), (name, lane, s)
        run.close()
        sol = op.solve(p, int(seed))
        # (the reference's solve mode leaves result.remaining_constraints at its
        # default and returns no solution unless it found a feasible one)
        assert sol["status"] == int(z[f"solve_{lane}"][0]), (name, lane)
        if sol["status"] == 0:
            assert np.array_equal(sol["x"], z[f"solve_x_{lane}"]), (name, lane)
            assert sol["value"] == float(z[f"solve_value_{lane}"]), (name, lane)
    op.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_engine_reproduces_reference_trajectory(name):
    from baryonyx_b200 import capi

    inst, p, z = util.load_golden(name)
    q = capi.Params()
    C.memmove(C.byref(q), C.byref(p), C.sizeof(q))
    seeds = [int(s) for s in z["seeds"]]
    pb = capi.Problem(inst)
    # one lane per batch: the recorded tail value differs per run
    for lane, seed in enumerate(seeds):
        bt = pb.batch(1, q)
        bt.set_tail_value(float(z[f"tail_{lane}"]))
        bt.begin([seed])
        assert np.array_equal(bt.state(0)["x"], z[f"x_{lane}"][0])
        for s, push, kappa in util.golden_kappa_schedule(z, p, inst["m"], lane):
            if push:
                rem = bt.sweep(p.pushing_k_factor * kappa, 1, p.pushing_objective_amplifier)
            else:
                rem = bt.sweep(kappa)
            st = bt.state(0)
            assert int(rem[0]) == int(z[f"rem_{lane}"][s + 1]), (name, lane, s)
            assert np.array_equal(st["x"], z[f"x_{lane}"][s + 1]), (name, lane, s)
            assert np.array_equal(st["pi"], z[f"pi_{lane}"][s + 1], equal_nan=True), (name, lane, s)
            assert np.array_equal(st["P"], z[f"P_{lane}"][s + 1], equal_nan=True), (name, lane, s)
            assert st["rng"] == int(z[f"rng_{lane}"][s + 1]), (name, lane, s)
        bt.begin([seed])
        out = bt.solve()
        assert int(out["status"][0]) == int(z[f"solve_{lane}"][0]), (name, lane)
        if int(out["status"][0]) == 0:
            assert np.array_equal(bt.best(0), z[f"solve_x_{lane}"]), (name, lane)
            assert out["value"][0] == float(z[f"solve_value_{lane}"]), (name, lane)
        bt.close()
    pb.close()
