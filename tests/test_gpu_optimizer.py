"""The optimizer on the GPU box (bxg_optimize_round): one lane = the
reference's optimize_functor at thread = 1, restart by restart, against

  * the committed traces of the COMPILED REFERENCE (tests/golden/optimizer_*.npz,
    written by tests/golden/make_golden_optimizer.py from bxref_opt_trace), and
  * the C restatement (ox_opt_trace_run, pinned to the same reference by
    tests/test_optimizer_trace.py) on more instances and seeds;

tolerance 0 on x, the engine state, kappa, pi, P and the whole population
(vectors, values, remaining constraints, hashes, loops, order).  Then several
lanes sharing one population: invariants, since the interleaving of inserts is
lock order there, as it is mutex order between the reference's threads."""
import ctypes as C
import os

import numpy as np
import pytest

from baryonyx_b200 import capi, synth
from oracle import pyoracle as po
from tests import util
from tests.optimizer_cases import CASES, TRACE_KEYS, case_params

pytestmark = pytest.mark.gpu


def engine_params(p):
    cp = po.clamp_params(p)
    q = capi.Params()
    C.memmove(C.byref(q), C.byref(cp), C.sizeof(q))
    return cp, q


def pop_arrays(pop, words):
    ms = pop.members()
    return {"pop_x": np.stack([m["x"] for m in ms]), "pop_value": np.array([m["value"] for m in ms]),
            "pop_remaining": np.array([m["remaining"] for m in ms], np.int32),
            "pop_hash": np.array([m["hash"] for m in ms], np.uint64),
            "pop_loop": np.array([m["loop"] for m in ms], np.int64)}


def run_against_trace(inst, q, seed, rounds, want, label):
    """Drives one lane through `rounds` restarts and compares with a trace."""
    pb = capi.Problem(inst, device=0)
    bt = pb.batch(1, q)
    pop = capi.Population(pb, q, seed, workers=1)
    pop.ctx(0, int(pop.generate_seeds(1)[0]))
    words = inst["n"] // 64 + 1
    sweeps_of = np.bincount(want["sweep_round"], minlength=rounds)
    updates_of = np.bincount(want["sweep_round"], weights=want["sweep_len"], minlength=rounds)
    for r in range(rounds + 1):
        kappa = pop.optimize_reinit(bt, first_round=(r == 0))
        st = bt.state(0)
        where = (label, "restart", r)
        assert kappa[0] == want["kappa"][r], where
        assert np.array_equal(st["x"], want["x"][r]), where
        assert st["rng"] == want["rng"][r], where
        assert np.array_equal(st["pi"], want["pi"][r]), where
        assert np.array_equal(st["P"], want["P"][r]), where
        got = pop_arrays(pop, words)
        for k in got:
            assert np.array_equal(got[k], want[k][r]), where + (k,)
        if r == rounds:
            break
        c0 = bt.counters()
        pop.optimize_run(bt)
        c1 = bt.counters()
        assert c1["sweeps"] - c0["sweeps"] == sweeps_of[r], where
        assert c1["updates"] - c0["updates"] == int(updates_of[r]), where
    pop.close()
    bt.close()
    pb.close()


@pytest.mark.parametrize("name", sorted(CASES))
def test_optimizer_lane_equals_reference_trace(name):
    factory, _, seed, rounds = CASES[name]
    want = np.load(os.path.join(util.GOLDEN_DIR, f"optimizer_{name}.npz"))
    _, q = engine_params(case_params(name, po))
    run_against_trace(factory(), q, seed, rounds, want, name)


ORACLE_CASES = [
    ("spp_ties", lambda: synth.spp(m=30, n=200, lo=3, hi=5, seed=9, cost_hi=3),
     dict(limit=200, pushes_limit=5, pushing_iteration_limit=8, init_population_size=30,
          init_kappa_improve_stop=0.04, kappa_step=0.01, init_mutation_variable_mean=0.02,
          init_mutation_variable_stddev=0.02), (1, 2, 3), 25),
    ("spp_bastert_insertion", lambda: synth.spp(m=20, n=120, lo=3, hi=5, seed=3, cost_hi=9),
     dict(limit=200, pushes_limit=3, pushing_iteration_limit=5, init_population_size=12,
          init_kappa_improve_stop=0.0, init_crossover_bastert_insertion=0.6, kappa_step=0.01), (4, 5, 6, 7), 30),
    ("scp_reversing", lambda: synth.scp(m=40, n=300, seed=12),
     dict(limit=50, pushes_limit=6, pushing_iteration_limit=4, init_population_size=16,
          init_kappa_improve_stop=0.02, order=1, init_mutation_variable_mean=0.01,
          init_mutation_variable_stddev=0.01), (8, 9), 40),
    ("mixed_101_no_mutation", lambda: synth.small_mixed(m=12, n=30, seed=8, family="101", maximize=True),
     dict(limit=80, pushes_limit=4, pushing_iteration_limit=6, init_population_size=10,
          init_kappa_improve_stop=0.02, init_mutation_value_mean=0.0, init_mutation_value_stddev=0.0), (3, 4), 30),
    ("telebus_Z", lambda: synth.telebus(m=30, n=100, seed=5),
     dict(limit=80, pushes_limit=4, pushing_iteration_limit=6, init_population_size=20,
          init_kappa_improve_stop=0.06, order=3), (5, 6), 25),
    ("spp_wide_mutation", lambda: synth.spp(m=50, n=5000, lo=3, hi=6, seed=4, cost_hi=20),
     dict(limit=30, pushes_limit=2, pushing_iteration_limit=3, init_population_size=10,
          init_kappa_improve_stop=0.0, init_mutation_variable_mean=0.3, init_mutation_variable_stddev=0.2), (1, 2), 12),
]


@pytest.mark.parametrize("name,make,over,seeds,rounds", ORACLE_CASES, ids=[c[0] for c in ORACLE_CASES])
def test_optimizer_lane_equals_oracle(name, make, over, seeds, rounds):
    inst = make()
    p = po.default_params()
    for k, v in over.items():
        setattr(p, k, v)
    cp, q = engine_params(p)
    op = po.OracleProblem(inst)
    for seed in seeds:
        want = op.optimize_trace(cp, seed, rounds)
        run_against_trace(inst, q, seed, rounds, want, f"{name}/{seed}")
    op.close()


def test_optimizer_lanes_share_one_population():
    """Eight workers on one population: every round the population stays
    ordered by (remaining constraints, objective), every member the lanes put
    there is what they say it is, and solutions arrive."""
    inst = synth.spp(m=20, n=120, lo=3, hi=5, seed=3, cost_hi=9)
    p = po.default_params()
    for k, v in dict(limit=400, pushes_limit=4, pushing_iteration_limit=6, init_population_size=20,
                     init_kappa_improve_stop=0.06, kappa_step=0.01).items():
        setattr(p, k, v)
    cp, q = engine_params(p)
    lanes = 8
    pb = capi.Problem(inst, device=0)
    bt = pb.batch(lanes, q)
    pop = capi.Population(pb, q, 77, workers=lanes)
    for l, s in enumerate(pop.generate_seeds(lanes)):
        pop.ctx(l, int(s))
    initial = {m["hash"] for m in pop.members()}
    for r in range(30):
        pop.optimize_round(bt, first_round=(r == 0))
        ms = pop.members()
        keys = [(m["remaining"], m["value"]) for m in ms]
        assert keys == sorted(keys), r
        xs = np.stack([m["x"] for m in ms])
        rem, val = pop.score(xs)
        for i, m in enumerate(ms):
            if m["hash"] in initial:
                continue  # crossover may have rewritten an initial member's bits without rescoring it
            assert pop.hash(m["x"]) == m["hash"], (r, i)
            assert val[i] == m["value"], (r, i)
            if m["remaining"] == 0:
                assert rem[i] == 0, (r, i)
    assert ms[0]["remaining"] == 0
    pop.close()
    bt.close()
    pb.close()
