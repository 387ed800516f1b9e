"""Optimizer population ops on the GPU box: the engine's
bxg_population_* (host streams + the device scoring kernel) against the C
oracle, which tests/test_population.py pins to the compiled reference."""
import ctypes as C

import numpy as np
import pytest

from baryonyx_b200 import capi, synth
from oracle import pyoracle as po
from tests import util
from tests.test_population import CASES, _plain, golden_population, population_script, same

pytestmark = pytest.mark.gpu


def _params(over):
    p = util.params(**over)
    cp = po.clamp_params(p)
    q = capi.Params()
    C.memmove(C.byref(q), C.byref(cp), C.sizeof(q))
    return cp, q


@pytest.mark.parametrize("name,make,over", CASES, ids=[c[0] for c in CASES])
def test_population_engine_vs_oracle(name, make, over):
    inst = make()
    cp, q = _params(over)
    op = po.OracleProblem(inst)
    pb = capi.Problem(inst, device=0)
    for master in (42, 7):
        a = po.Population(op, master, threads=2, params=cp)
        b = capi.Population(pb, q, master, workers=2)
        ra = population_script(a, inst["n"], cp.kappa_min, cp.kappa_max, 40, 11)
        rb = population_script(b, inst["n"], cp.kappa_min, cp.kappa_max, 40, 11)
        assert ra["ctor_rng"] == rb["ctor_rng"]
        assert same(ra["members0"], rb["members0"]), "population after the constructor"
        for i, (ta, tb) in enumerate(zip(ra["trace"], rb["trace"])):
            assert same(ta, tb), (i, ta[0])
        assert same(ra["members1"], rb["members1"])
        a.close()
        b.close()
    pb.close()
    op.close()


def test_score_vectors_matches_members():
    """The device scoring kernel on arbitrary vectors, BASELINE-sized SPP."""
    inst = synth.spp()
    cp, q = _params({})
    pb = capi.Problem(inst, device=0)
    op = po.OracleProblem(inst)
    a = po.Population(op, 3, threads=1, params=cp)
    b = capi.Population(pb, q, 3, workers=1)
    ma, mb = a.members(), b.members()
    assert same(ma, mb)
    xs = np.stack([m["x"] for m in mb])
    rem, val = b.score(xs)
    assert rem.tolist() == [m["remaining"] for m in ma]
    assert val.tolist() == [m["value"] for m in ma]
    assert b.hash(xs[0]) == ma[0]["hash"]
    a.close()
    b.close()
    pb.close()
    op.close()


ROUND_CASES = [
    # never feasible within the limit: every call ends in try_advance
    ("advance", lambda: synth.spp(m=40, n=300, lo=3, hi=6, seed=2, cost_hi=5), {"limit": 60}, False),
    # some lanes reach feasibility: push phase, try_update
    ("update", lambda: synth.small_mixed(m=14, n=40, seed=21, family="01", equality=True, cost_hi=5),
     {"limit": 150, "kappa_step": 0.01}, True),
]


@pytest.mark.parametrize("name,make,extra,expect_solved", ROUND_CASES, ids=[c[0] for c in ROUND_CASES])
def test_optimize_rounds_engine_vs_oracle(name, make, extra, expect_solved):
    """bxg_optimize_round (reinit -> device run with the optimizer's kappa
    handling -> try_update / try_advance) against the same schedule composed
    from the oracle's pieces, several rounds, Solver state carried over."""
    inst = make()
    over = {"pushes_limit": 4, "pushing_iteration_limit": 6, "init_population_size": 12,
            "init_kappa_improve_stop": 0.04, "init_mutation_variable_mean": 0.05,
            "init_mutation_variable_stddev": 0.02}
    over.update(extra)
    cp, q = _params(over)
    lanes, master, rounds = 6, 99, 5
    from baryonyx_b200 import exchange

    seeds = exchange.lane_seeds(master, lanes)

    # --- oracle side
    op = po.OracleProblem(inst)
    opop = po.Population(op, master, threads=lanes, params=cp)
    runs = []
    for l in range(lanes):
        opop.ctx(l, int(seeds[l]))
        runs.append(op.run(cp, 1))
    words = inst["n"] // 64 + 1
    xs = [np.zeros(words, np.uint64) for _ in range(lanes)]

    # --- engine side
    pb = capi.Problem(inst, device=0)
    bt = pb.batch(lanes, q)
    pop = capi.Population(pb, q, master, workers=lanes)
    for l in range(lanes):
        pop.ctx(l, int(seeds[l]))

    solved = 0
    for rnd in range(rounds):
        out = pop.optimize_round(bt, first_round=(rnd == 0))
        results = []
        for l in range(lanes):
            kappa = opop.reinit(l, cp.kappa_min, cp.kappa_max, xs[l])
            r = runs[l]
            r.set_rng(opop.ctx_rng(l))
            r.set_x(xs[l])
            res = r.solve(kappa, optimizer=True)
            st = r.state()
            xs[l] = st["x"].copy()
            results.append((res, st))
        for l, (res, st) in enumerate(results):
            # the worker's engine continues from where the run left it
            opop.set_ctx_rng(l, st["rng"])
            if res["remaining"] == 0:
                opop.try_update(l, res["best_x"], res["value"], res["loop"])
                solved += 1
            else:
                opop.try_advance(l, xs[l], res["remaining"], cp.limit)
            assert out["status"][l] == res["status"], (rnd, l)
            assert out["remaining"][l] == res["remaining"], (rnd, l)
            if res["remaining"] == 0:
                assert out["value"][l] == res["value"], (rnd, l)
            es = bt.state(l)
            for k in ("x", "pi", "P"):
                assert np.array_equal(es[k], st[k]), (rnd, l, k)
            assert pop.ctx_rng(l) == opop.ctx_rng(l), (rnd, l)
        assert same(pop.members(), opop.members()), rnd
    assert (solved > 0) == expect_solved
    for r in runs:
        r.close()
    opop.close()
    op.close()
    pop.close()
    bt.close()
    pb.close()


@pytest.mark.parametrize("name,make,over", CASES, ids=[c[0] for c in CASES])
def test_population_engine_vs_golden(name, make, over):
    """bxg_population_* against the committed traces of the compiled reference."""
    g = golden_population(name)
    inst = make()
    cp, q = _params(over)
    pb = capi.Problem(inst, device=0)
    for master, want in g["masters"].items():
        pop = capi.Population(pb, q, int(master), workers=2)
        got = _plain(population_script(pop, inst["n"], cp.kappa_min, cp.kappa_max, g["steps"], g["data_seed"]))
        assert got == want, (name, master)
        pop.close()
    pb.close()
