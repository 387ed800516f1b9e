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
