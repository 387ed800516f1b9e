"""The BASELINE.json configurations at their FULL sizes:
the engine against the oracle, state by state, for as many sweeps as the
oracle finishes in seconds, plus properties that do not need the oracle."""
import numpy as np
import pytest

from baryonyx_b200 import capi, synth
from oracle import pyoracle as po
from tests import util
from tests.test_gpu_parity import lockstep_engine, solve_case, to_capi

pytestmark = pytest.mark.gpu


def violated_rows(inst, x_words):
    """Number of constraints x violates, straight from the CSR arrays."""
    bits = np.unpackbits(np.ascontiguousarray(x_words).view(np.uint8), bitorder="little")
    v = bits[inst["col"]].astype(np.int64) * inst["factor"]
    act = np.add.reduceat(v, inst["row_ptr"][:-1])
    return int(np.count_nonzero((act < inst["cmin"]) | (act > inst["cmax"])))


def test_c2_spp_500x10000():
    inst = synth.spp()
    assert (inst["m"], inst["n"]) == (500, 10000)
    lockstep_engine(inst, util.params(), [42, 43, 44, 45], 12, pushes_every=5)
    solve_case(inst, util.params(limit=25, pushes_limit=2, pushing_iteration_limit=3), [7, 8])


def test_c3_csplib022_30x50000_101():
    inst = synth.csplib022()
    assert (inst["m"], inst["n"]) == (30, 50000)
    assert inst["problem_type"] == 1  # equalities_101
    lockstep_engine(inst, util.params(kappa_step=0.01), [1, 2], 10, pushes_every=4)


@pytest.mark.parametrize("order", ["infeasibility-decr", "infeasibility-incr"])
def test_c4_telebus_1500x7000_z(order):
    inst = synth.telebus()
    assert (inst["m"], inst["n"]) == (1500, 7000)
    lockstep_engine(inst, util.params(order=order, kappa_step=0.005, w=1), [11, 12], 6, pushes_every=3)


def test_c5_rail_4284x1092610():
    inst = synth.rail()
    assert (inst["m"], inst["n"]) == (4284, 1092610)
    # x = 0 to start with (bastert on positive costs, no random flips): every
    # covering row is violated and the first sweep updates all 4284 of them
    p = util.params(init_policy_random=0.0)
    lockstep_engine(inst, p, [5, 6], 3)


def test_properties_at_full_size():
    """remaining == violated rows of x recomputed on the host; the scoring
    kernel agrees; equal seeds give equal lanes; a run is repeatable."""
    inst = synth.spp()
    cp = po.clamp_params(util.params(limit=40))
    q = to_capi(cp)
    pb = capi.Problem(inst)
    bt = pb.batch(64, q)
    seeds = np.arange(64, dtype=np.uint32) // 2 + 100  # pairs of equal seeds
    bt.begin(seeds)
    kappa = 0.0
    for s in range(15):
        rem = bt.sweep(kappa)
        kappa += 0.002
    states = [bt.state(l) for l in range(64)]
    for l in range(0, 64, 2):
        ok, k = util.states_equal(states[l], states[l + 1])
        assert ok, (l, k)
    for l in range(0, 64, 8):
        assert rem[l] == violated_rows(inst, states[l]["x"]), l
    pop = capi.Population(pb, q, 1, workers=1)
    r2, _ = pop.score(np.stack([states[l]["x"] for l in range(64)]))
    assert r2.tolist() == list(rem)
    pop.close()

    first = bt.solve()
    xs = [bt.best(l).copy() for l in range(4)]
    bt.begin(seeds)
    for s in range(15):
        bt.sweep(0.002 * s)
    again = bt.solve()
    for k in first:
        assert np.array_equal(first[k], again[k]), k
    for l in range(4):
        assert np.array_equal(xs[l], bt.best(l))
    bt.close()
    pb.close()
