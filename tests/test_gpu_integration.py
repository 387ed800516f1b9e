"""The drop-in boundary on the reference's own types:
integration/itm_b200.cpp -- reference problem -> make_merged_constraints ->
C ABI -> reference result -- compiled against the reference's headers into
oracle/_ref/libbxref.so, against the reference's own solve_problem<> on the
same LP text, parameters and seed.  Needs the prebuilt harness library (it
travels to the GPU box) and a GPU."""
import os

import numpy as np
import pytest

from baryonyx_b200 import synth
from oracle import pyoracle as po
from tests import util

pytestmark = [pytest.mark.gpu, pytest.mark.ref]

CASES = [
    ("eq01", lambda: synth.spp(m=40, n=300, lo=3, hi=6, seed=2, cost_hi=5), dict(limit=80)),
    ("ineq01", lambda: synth.scp(m=50, n=250, rows_per_col=4, seed=1), dict(limit=80)),
    ("eq01_solved", lambda: synth.small_mixed(m=14, n=40, seed=21, family="01", equality=True, cost_hi=5),
     dict(limit=150, kappa_step=0.01)),
    ("ineq101", lambda: synth.small_mixed(m=14, n=40, seed=22, family="101", equality=False, cost_hi=5),
     dict(limit=150, kappa_step=0.01)),
    ("ineqZ", lambda: synth.telebus(m=90, n=400, seed=4), dict(limit=60, order="infeasibility-decr")),
]


@pytest.mark.parametrize("name,make,over", CASES, ids=[c[0] for c in CASES])
def test_reference_api_over_the_engine(name, make, over):
    inst = make()
    p = util.params(pushes_limit=3, pushing_iteration_limit=5, **over)
    lp = util.lp_of(inst)
    try:
        # the release flavour of the reference (contract checks compiled out, its
        # WITH_FULL_OPTIMIZATION option): the checked build aborts on some Z rows
        # (itm-solver-inequalities-Z.cpp:436) where the release build carries on
        rp = po.RefProblem(lp, p, fast=True)
    finally:
        os.unlink(lp)
    for seed in (11, 12, 13):
        ref = rp.solve(seed)
        eng = rp.solve_b200(seed, restarts=1)
        # the whole baryonyx::result the two return: status, the solution's
        # variables and value (+inf while no feasible vector was found), and the
        # fields the reference leaves at their defaults
        assert eng["status"] == ref["status"], (seed, eng["status"], ref["status"])
        assert np.array_equal(eng["x"], ref["x"]), seed
        assert eng["value"] == ref["value"], seed
        assert eng["remaining"] == ref["remaining"] and eng["loop"] == 0, seed
    # several restarts: the best lane is returned, never worse than lane 0
    one = rp.solve_b200(11, restarts=1)
    many = rp.solve_b200(11, restarts=8)
    assert many["status"] == 0 or one["status"] != 0
    if one["status"] == 0:
        better = many["value"] >= one["value"] if inst["maximize"] else many["value"] <= one["value"]
        assert better
    rp.close()


@pytest.mark.parametrize("name", ["spp_eq01_feasible", "scp_ineq01", "mixed_ineq101_max", "telebus_ineqZ_b"])
def test_reference_optimize_over_the_engine(name):
    """baryonyx::itm::optimize_b200 with one worker, bounded by restarts, against
    the reference's own optimize_functor at thread = 1 (bxref_opt_trace): the
    result's solution is the best member of the reference's population after the
    same number of restarts -- same remaining constraints, same value, same x."""
    from tests.optimizer_cases import CASES as OPT, case_params

    factory, _, seed, rounds = OPT[name]
    p = case_params(name, po)
    lp = util.lp_of(factory())
    try:
        rp = po.RefProblem(lp, p, fast=True)
    finally:
        os.unlink(lp)
    tr = rp.optimize_trace(seed, rounds, p.init_population_size)
    eng = rp.optimize_b200(seed, workers=1, max_rounds=rounds)
    want_rem, want_val = int(tr["pop_remaining"][rounds][0]), float(tr["pop_value"][rounds][0])
    assert eng["remaining"] == want_rem
    assert eng["value"] == want_val
    bits = np.unpackbits(tr["pop_x"][rounds][0].view(np.uint8), bitorder="little")[: rp.n]
    assert np.array_equal(eng["x"], bits.astype(np.int8))
    assert eng["status"] == (0 if want_rem == 0 else 3)  # success / time_limit_reached, :873-878
    assert eng["calls"]["start"] == 1 and eng["calls"]["finish"] == 1 and eng["solutions"] == 1
    rp.close()


def test_reference_optimize_callbacks_and_time_limit():
    """time_limit drives the run (the kernels see the deadline too), ctx.update is
    fed about once a second, storage_type five returns five solutions."""
    import time

    inst = synth.spp(m=40, n=300, lo=3, hi=6, seed=2, cost_hi=5)
    p = util.params(limit=200, pushes_limit=4, pushing_iteration_limit=6, init_population_size=20)
    lp = util.lp_of(inst)
    try:
        rp = po.RefProblem(lp, p, fast=True)
    finally:
        os.unlink(lp)
    t0 = time.time()
    eng = rp.optimize_b200(7, workers=64, time_limit=2.5, storage=2)
    dt = time.time() - t0
    assert 2.4 <= dt < 6.0, dt
    assert eng["calls"] == {"start": 1, "update": eng["calls"]["update"], "finish": 1}
    assert eng["calls"]["update"] >= 1
    assert eng["solutions"] == 5
    assert eng["status"] in (0, 3)
    rp.close()
