"""Optimizer population ops: the C restatement against the
compiled reference (best_solution_recorder / storage / local_context of
lib/src/itm-optimizer-common.hpp:52-594), op by op on the same seeds."""
import os

import numpy as np
import pytest

from baryonyx_b200 import synth
from oracle import pyoracle as po
from tests import util


def population_script(pop, n, kappa_min, kappa_max, steps, data_seed, threads=2):
    """A fixed sequence of worker calls; returns everything observable."""
    g = np.random.default_rng(data_seed)
    words = n // 64 + 1
    out = {"ctor_rng": pop.rng_after_ctor, "members0": pop.members(), "trace": []}
    xs = []
    for t in range(threads):
        pop.ctx(t, 1000 + 17 * t)
        xs.append(pop.member(t)["x"].copy())
    for s in range(steps):
        t = s % threads
        x = xs[t]
        kappa = pop.reinit(t, kappa_min, kappa_max, x)
        out["trace"].append(("reinit", kappa, x.copy(), pop.ctx_rng(t)))
        # a worker's result: now and then a solution, otherwise an advance
        y = x.copy()
        flips = g.integers(0, n, 3)
        for f in flips:
            y[f // 64] ^= np.uint64(1) << np.uint64(f % 64)
        if s % 3 == 0:
            pop.try_update(t, y, float(g.integers(1, 50)), int(s))
        else:
            pop.try_advance(t, y, int(g.integers(1, 6)), 1000)
        if s % 5 == 4:  # the same vector again: must be refused without touching the engine
            pop.try_advance(t, y, int(g.integers(1, 6)), 1000)
        out["trace"].append(("insert", pop.ctx_rng(t), [m["hash"] for m in pop.members()],
                             [m["remaining"] for m in pop.members()], [m["value"] for m in pop.members()]))
        xs[t] = y
    z = xs[0].copy()
    pop.crossover(0, z)
    out["trace"].append(("crossover", z.copy(), pop.ctx_rng(0)))
    pop.mutation(0, z)
    out["trace"].append(("mutation", z.copy(), pop.ctx_rng(0)))
    out["members1"] = pop.members()
    return out


def same(a, b):
    if isinstance(a, dict):
        return a.keys() == b.keys() and all(same(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
    if isinstance(a, np.ndarray):
        return np.array_equal(a, b)
    if isinstance(a, float):
        return a == b or (a != a and b != b)
    return a == b


CASES = [
    # name, instance, parameter overrides
    ("spp", lambda: synth.spp(m=30, n=200, lo=2, hi=5, seed=3, cost_hi=9), {}),
    ("scp_ineq", lambda: synth.scp(m=25, n=150, seed=4), {"init_population_size": 21}),
    ("spp_mut", lambda: synth.spp(m=20, n=130, lo=2, hi=4, seed=5, cost_hi=4),
     {"init_mutation_variable_mean": 0.2, "init_mutation_variable_stddev": 0.1,
      "init_crossover_bastert_insertion": 0.6, "init_kappa_improve_stop": 0.04,
      "init_population_size": 10}),
]


@pytest.mark.ref
@pytest.mark.parametrize("name,make,over", CASES, ids=[c[0] for c in CASES])
def test_population_oracle_vs_reference(name, make, over):
    inst = make()
    p = util.params(**over)
    lp = util.lp_of(inst)
    try:
        rp = po.RefProblem(lp, p)
    finally:
        os.unlink(lp)
    cp = rp.params()
    ri = rp.instance()
    op = po.OracleProblem(ri)
    for master in (42, 7):
        a = po.Population(rp, master, threads=2)
        b = po.Population(op, master, threads=2, params=cp)
        ra = population_script(a, rp.n, cp.kappa_min, cp.kappa_max, 40, 11)
        rb = population_script(b, rp.n, cp.kappa_min, cp.kappa_max, 40, 11)
        assert ra["ctor_rng"] == rb["ctor_rng"]
        assert same(ra["members0"], rb["members0"]), "population after the constructor"
        for i, (ta, tb) in enumerate(zip(ra["trace"], rb["trace"])):
            assert same(ta, tb), (i, ta[0])
        assert same(ra["members1"], rb["members1"])
        a.close()
        b.close()
    op.close()
    rp.close()


@pytest.mark.ref
def test_bit_array_hash_matches_reference():
    R = po.ref()
    O = po.oracle()
    po._bind_population(R, "bxref_pop_", "bxref_bit_array_hash")
    po._bind_population(O, "ox_pop_", "ox_bit_array_hash")
    g = np.random.default_rng(0)
    for n in (1, 63, 64, 65, 1000):
        x = g.integers(0, 2**63, n // 64 + 1, dtype=np.uint64)
        assert R.bxref_bit_array_hash(n, x) == O.ox_bit_array_hash(n, x)


# ---- golden traces of the compiled reference (tests/golden/make_golden_population.py)
def _plain(o):
    from tests.golden.make_golden_population import plain

    return plain(o)


def golden_population(name):
    import gzip
    import json

    path = os.path.join(os.path.dirname(__file__), "golden", f"population_{name}.json.gz")
    with gzip.open(path, "rb") as f:
        return json.loads(f.read().decode())


@pytest.mark.parametrize("name,make,over", CASES, ids=[c[0] for c in CASES])
def test_population_oracle_vs_golden(name, make, over):
    """The C restatement against the committed traces of the reference (runs
    where /root/reference does not exist)."""
    g = golden_population(name)
    inst = make()
    cp = po.clamp_params(util.params(**over))
    op = po.OracleProblem(inst)
    for master, want in g["masters"].items():
        pop = po.Population(op, int(master), threads=2, params=cp)
        got = _plain(population_script(pop, inst["n"], cp.kappa_min, cp.kappa_max, g["steps"], g["data_seed"]))
        assert got == want, (name, master)
        pop.close()
    op.close()
