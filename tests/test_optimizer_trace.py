"""The optimizer loop at thread = 1 (optimize_functor::operator(),
lib/src/itm-optimizer-common.hpp:596-752): the C restatement (ox_opt_trace_run)
against the compiled reference's own functor (bxref_opt_trace), restart by
restart -- x, engine, kappa, pi, P, the whole population, and the sweep log --
and both against the committed golden traces."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from tests import util
from tests.optimizer_cases import CASES, TRACE_KEYS, case_params


def oracle_trace(name, inst=None, cp=None):
    factory, _, seed, rounds = CASES[name]
    inst = inst or factory()
    cp = cp or po.clamp_params(case_params(name, po))
    op = po.OracleProblem(inst)
    try:
        return op.optimize_trace(cp, seed, rounds)
    finally:
        op.close()


@pytest.mark.ref
@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_trace_equals_reference(name):
    factory, _, seed, rounds = CASES[name]
    inst = factory()
    p = case_params(name, po)
    path = util.lp_of(inst)
    # the Z solver's own contract check (itm-solver-inequalities-Z.cpp:436) fires
    # on some restarts: the build without contract checks has the same arithmetic
    pb = po.RefProblem(path, p, fast=True)
    os.unlink(path)
    try:
        ref = pb.optimize_trace(seed, rounds, p.init_population_size)
        got = oracle_trace(name, pb.instance(), pb.params())
    finally:
        pb.close()
    assert ref["x"].shape[0] == rounds + 1
    for k in TRACE_KEYS:
        assert np.array_equal(ref[k], got[k]), k


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_trace_equals_golden(name):
    path = os.path.join(util.GOLDEN_DIR, f"optimizer_{name}.npz")
    gold = np.load(path)
    got = oracle_trace(name)
    for k in TRACE_KEYS:
        assert np.array_equal(gold[k], got[k]), k
