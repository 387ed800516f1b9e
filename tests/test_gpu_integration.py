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
