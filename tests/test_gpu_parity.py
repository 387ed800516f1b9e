"""Parity of the CUDA engine (through the C ABI) against the oracle on the
same seeded inputs: bit-exact x, order lists, random-engine state, and pi / P
(bit-exact too: fp64 deterministic mode, no FMA on either side)."""
import ctypes as C

import numpy as np
import pytest

from baryonyx_b200 import capi, synth
from oracle import pyoracle as po
from tests import util

pytestmark = pytest.mark.gpu


def to_capi(p: po.Params) -> capi.Params:
    q = capi.Params()
    C.memmove(C.byref(q), C.byref(p), C.sizeof(q))
    return q


def lockstep_engine(inst, p, seeds, sweeps, pushes_every=0, tail=0.0):
    """Engine lanes vs one oracle run per lane, compared after every sweep."""
    cp = po.clamp_params(p)
    op = po.OracleProblem(inst)
    op.set_tail_value(tail)
    runs = [op.run(cp, s) for s in seeds]
    pb = capi.Problem(inst)
    bt = pb.batch(len(seeds), to_capi(cp))
    bt.set_tail_value(tail)
    bt.begin(seeds)

    info = bt.info()
    assert (info["delta"], info["theta"]) == runs[0].info()
    assert np.array_equal(bt.norm_cost(), runs[0].norm_cost())

    infeas = cp.order in (3, 4)

    def compare(tag):
        for l, r in enumerate(runs):
            a, b = r.state(), bt.state(l)
            if infeas:
                a["order_len"] *= 2
                b["order_len"] *= 2
            ok, k = util.states_equal(a, b)
            assert ok, f"lane {l} ({tag}): '{k}' differs"

    compare("initial")
    m = inst["m"]
    kappa = np.full(len(seeds), cp.kappa_min)
    total_updates = 0
    for s in range(sweeps):
        push = 1 if (pushes_every and s % pushes_every == pushes_every - 1) else 0
        if push:
            rem = bt.sweep(cp.pushing_k_factor * kappa, 1, cp.pushing_objective_amplifier)
            ref = [r.sweep(cp.pushing_k_factor * kappa[l], 1, cp.pushing_objective_amplifier)
                   for l, r in enumerate(runs)]
        else:
            rem = bt.sweep(kappa)
            ref = [r.sweep(kappa[l]) for l, r in enumerate(runs)]
        assert list(rem) == ref, f"remaining differs at sweep {s}"
        compare(f"sweep {s}")
        if s > cp.w:
            kappa = kappa + cp.kappa_step * (np.asarray(ref) / m) ** cp.alpha
        kappa = np.minimum(kappa, 0.999)
    obj = bt.objective()
    for l, r in enumerate(runs):
        assert obj[l] == r.objective()
        total_updates += r.updates()
    assert bt.counters()["updates"] == total_updates
    for r in runs:
        r.close()
    bt.close(), pb.close(), op.close()


@pytest.mark.parametrize("family", ["01", "101", "Z"])
@pytest.mark.parametrize("equality", [False, True])
@pytest.mark.parametrize("maximize", [False, True])
def test_sweeps_small(family, equality, maximize):
    inst = synth.small_mixed(m=14, n=36, seed=11, family=family, equality=equality, maximize=maximize,
                             cost_hi=4)
    p = util.params(limit=200, kappa_step=0.01, w=3)
    lockstep_engine(inst, p, [41, 42, 43, 44], 50, pushes_every=7)


@pytest.mark.parametrize("order", ["none", "reversing", "random-sorting", "infeasibility-decr",
                                   "infeasibility-incr", "lagrangian-decr", "lagrangian-incr",
                                   "pi-sign-change", "cycle"])
@pytest.mark.parametrize("family", ["01", "101", "Z"])
def test_sweeps_orders(order, family):
    if order in ("pi-sign-change", "cycle"):
        inst = synth.small_mixed(m=30, n=20, seed=5, family=family, cost_hi=3, density=0.3)
    else:
        inst = synth.small_mixed(m=20, n=50, seed=6, family=family, cost_hi=3, density=0.2)
    p = util.params(limit=100, kappa_step=0.01, w=2, order=order)
    lockstep_engine(inst, p, [7, 8, 9], 30, pushes_every=5)


@pytest.mark.parametrize("init_policy", [0, 1, 2])
@pytest.mark.parametrize("cost_norm", [0, 2, 3, 4])
def test_sweeps_init_and_norm(init_policy, cost_norm):
    inst = synth.small_mixed(m=16, n=40, seed=9, family="101", cost_hi=5)
    p = util.params(init_policy=init_policy, cost_norm=cost_norm, init_policy_random=0.3, kappa_step=0.01)
    lockstep_engine(inst, p, [3, 4], 25)


def test_sweeps_theta_delta_alpha():
    inst = synth.small_mixed(m=16, n=40, seed=2, family="01", cost_hi=6)
    for theta, delta, alpha in ((0.0, -1, 1.0), (1.0, 0.0, 2.0), (0.7, 0.01, 0.5), (0.5, -1, 0.0)):
        p = util.params(theta=theta, delta=delta, alpha=alpha, kappa_step=0.02, w=1)
        lockstep_engine(inst, p, [5, 6], 25)


def test_sweeps_long_rows_with_ties():
    # rows longer than 16 elements: introsort proper, ties everywhere
    inst = synth.scp(m=40, n=400, rows_per_col=6, seed=3, cost_hi=2)
    p = util.params(kappa_step=0.005, init_policy_random=0.0)
    lockstep_engine(inst, p, [11, 12], 30, pushes_every=9)
    lockstep_engine(inst, p, [11], 10, tail=1.0)
    inst = synth.spp(m=60, n=600, lo=3, hi=6, seed=4, cost_hi=3)
    p = util.params(kappa_step=0.002)
    lockstep_engine(inst, p, [12, 13, 14], 60)


def test_sweeps_scp_spp_shaped_medium():
    inst = synth.scp(m=200, n=1000, rows_per_col=4, seed=1)
    p = util.params(init_policy_random=0.5)
    lockstep_engine(inst, p, [42, 43], 12, pushes_every=4)
    inst = synth.spp(m=120, n=2000, lo=4, hi=7, seed=2)
    lockstep_engine(inst, util.params(), [42, 43, 44, 45], 40)


def solve_case(inst, p, seeds, tail=0.0):
    cp = po.clamp_params(p)
    op = po.OracleProblem(inst)
    op.set_tail_value(tail)
    pb = capi.Problem(inst)
    bt = pb.batch(len(seeds), to_capi(cp))
    bt.set_tail_value(tail)
    bt.begin(seeds)
    out = bt.solve()
    upd = 0
    swp = 0
    for l, s in enumerate(seeds):
        o = op.solve(cp, s)
        assert out["status"][l] == o["status"], (l, out["status"][l], o["status"])
        assert np.array_equal(bt.best(l), o["x"]), l
        assert out["value"][l] == o["value"], l
        assert out["remaining"][l] == o["remaining"], l
        assert out["loop"][l] == o["loop"], l
        upd += o["updates"]
        swp += o["sweeps"]
    c = bt.counters()
    assert c["updates"] == upd
    assert c["sweeps"] == swp
    bt.close(), pb.close(), op.close()


@pytest.mark.parametrize("family", ["01", "101", "Z"])
def test_solve_whole(family):
    for equality in (False, True):
        inst = synth.small_mixed(m=14, n=40, seed=21, family=family, equality=equality, cost_hi=5)
        p = util.params(limit=300, pushes_limit=10, pushing_iteration_limit=8, kappa_step=0.005)
        solve_case(inst, p, [1, 2, 3, 4, 5, 6, 7, 8])


def test_solve_scp_spp_shaped():
    p = util.params(limit=200, pushes_limit=5, pushing_iteration_limit=10)
    solve_case(synth.scp(m=50, n=250, rows_per_col=4, seed=1), p, [42, 43, 44])
    solve_case(synth.spp(m=60, n=500, lo=3, hi=6, seed=2), p, [42, 43, 44])
    # infeasible within the limit: status limit_reached / kappa_max_reached paths
    p = util.params(limit=15, pushes_limit=2, pushing_iteration_limit=3)
    solve_case(synth.spp(m=60, n=500, lo=3, hi=6, seed=2), p, [1, 2])
    p = util.params(limit=400, kappa_step=0.05, kappa_max=0.3, w=0)
    solve_case(synth.spp(m=60, n=500, lo=3, hi=6, seed=2), p, [1, 2])


def test_sweeps_z_sub_solvers():
    """Z rows of every kind: linear, exhaustive (< 10 elements), branch and
    bound (10..16 elements), constrained and unconstrained, and a row long
    enough for the CTA-wide ordering (> 32 elements)."""
    inst = synth.telebus(m=90, n=400, seed=4)
    for order in ("none", "infeasibility-decr", "infeasibility-incr"):
        p = util.params(limit=100, kappa_step=0.002, w=2, order=order)
        lockstep_engine(inst, p, [21, 22, 23], 40, pushes_every=6)
    inst = synth.telebus(m=40, n=300, seed=6, linear_hi=90)
    lockstep_engine(inst, util.params(kappa_step=0.01), [5, 6], 25, pushes_every=5)
    inst = synth.small_mixed(m=12, n=60, seed=14, family="Z", density=0.2, cost_hi=2, maximize=True)
    lockstep_engine(inst, util.params(kappa_step=0.01), [5, 6], 25, pushes_every=5)


def test_solve_telebus_shaped():
    inst = synth.telebus(m=150, n=700, seed=4)
    p = util.params(limit=150, pushes_limit=3, pushing_iteration_limit=5, order="infeasibility-decr")
    solve_case(inst, p, [1, 2, 3, 4])


def test_set_state_roundtrip_and_carry_over():
    """bxg_batch_set_state taps + a second run that keeps P / pi (optimizer
    workers never reset their solver between restarts)."""
    inst = synth.small_mixed(m=14, n=36, seed=3, family="01", cost_hi=4)
    cp = po.clamp_params(util.params(kappa_step=0.01))
    op = po.OracleProblem(inst)
    r = op.run(cp, 5)
    pb = capi.Problem(inst)
    bt = pb.batch(2, to_capi(cp))
    bt.begin([5, 6])
    for s in range(5):
        bt.sweep(0.01)
        r.sweep(0.01)
    st = r.state()
    # move lane 1 onto lane 0's trajectory through the taps
    bt.set_state(1, x_words=st["x"], pi=st["pi"], P=st["P"], rng=st["rng"])
    for s in range(5):
        rem = bt.sweep(0.02)
        ref = r.sweep(0.02)
        assert rem[0] == ref and rem[1] == ref
    a, b0, b1 = r.state(), bt.state(0), bt.state(1)
    assert util.states_equal(a, b0)[0] and util.states_equal(a, b1)[0]
    # new x, same solver state
    x2 = st["x"].copy()
    x2[0] ^= np.uint64(0x5555)
    r.set_x(x2)
    bt.set_state(0, x_words=x2)
    bt.set_state(1, x_words=x2)
    for s in range(5):
        rem = bt.sweep(0.03)
        ref = r.sweep(0.03)
        assert rem[0] == ref and rem[1] == ref
    assert util.states_equal(r.state(), bt.state(0))[0]
    r.close(), bt.close(), pb.close(), op.close()


def test_parameter_sweep_lanes():
    """bxg_batch_set_lane_params: every lane runs its own theta / delta /
    kappa_min / kappa_step / init_policy_random (the grid of the reference's
    manual auto-tuner, manual-optimizer.cpp:130-160) and matches an oracle run
    with those parameters, sweep by sweep and as a whole solve."""
    inst = synth.spp(m=60, n=600, lo=3, hi=6, seed=4, cost_hi=3)
    grid = [
        dict(theta=0.5, delta=-1.0, kappa_min=0.0, kappa_step=1e-3, init_policy_random=0.5),
        dict(theta=0.3, delta=-1.0, kappa_min=0.05, kappa_step=2e-3, init_policy_random=0.1),
        dict(theta=0.7, delta=0.01, kappa_min=0.1, kappa_step=5e-3, init_policy_random=0.9),
        dict(theta=1.5, delta=-3.0, kappa_min=0.0, kappa_step=1e-2, init_policy_random=0.0),  # clamped
        dict(theta=0.0, delta=0.0, kappa_min=0.2, kappa_step=0.0, init_policy_random=1.0),
    ]
    base = dict(limit=60, pushes_limit=3, pushing_iteration_limit=5)
    seeds = [31, 32, 33, 34, 35]
    cps = [po.clamp_params(util.params(**base, **g)) for g in grid]
    op = po.OracleProblem(inst)
    pb = capi.Problem(inst)
    bt = pb.batch(len(grid), to_capi(po.clamp_params(util.params(**base))))
    bt.set_lane_params(**{k: np.array([g[k] for g in grid]) for k in grid[0]})

    # sweep by sweep with each lane's kappa schedule
    bt.begin(seeds)
    runs = [op.run(cp, s) for cp, s in zip(cps, seeds)]
    kappa = np.array([cp.kappa_min for cp in cps])
    for s in range(25):
        rem = bt.sweep(kappa)
        ref = [r.sweep(kappa[l]) for l, r in enumerate(runs)]
        assert list(rem) == ref, s
        for l, r in enumerate(runs):
            ok, k = util.states_equal(r.state(), bt.state(l))
            assert ok, (s, l, k)
        if s > cps[0].w:
            kappa = kappa + np.array([cp.kappa_step for cp in cps]) * (np.asarray(ref) / inst["m"]) ** cps[0].alpha
    for r in runs:
        r.close()

    # whole solves
    bt.begin(seeds)
    out = bt.solve()
    for l, (cp, s) in enumerate(zip(cps, seeds)):
        o = op.solve(cp, s)
        assert out["status"][l] == o["status"], l
        assert out["remaining"][l] == o["remaining"], l
        assert out["loop"][l] == o["loop"], l
        assert out["value"][l] == o["value"], l
        assert np.array_equal(bt.best(l), o["x"]), l

    # back to the batch's values
    bt.set_lane_params()
    bt.begin(seeds)
    out = bt.solve()
    o = op.solve(po.clamp_params(util.params(**base)), seeds[3])
    assert out["loop"][3] == o["loop"] and np.array_equal(bt.best(3), o["x"])
    bt.close(), pb.close(), op.close()


def test_manual_grid_lanes():
    """bxg_manual_grid -> bxg_batch_set_lane_params: a slice of the manual
    auto-tuner's 3125 parameter sets (manual-optimizer.cpp:113-160) on the
    lanes of one batch; every lane's whole solve equals the oracle's with that
    lane's parameters."""
    inst = synth.spp(m=60, n=600, lo=3, hi=6, seed=4, cost_hi=3)
    base = dict(limit=40, pushes_limit=2, pushing_iteration_limit=4, theta=0.6, delta=0.015,
                kappa_min=0.01, kappa_step=2e-3, init_policy_random=0.25)
    q = to_capi(po.clamp_params(util.params(**base)))
    grid = capi.manual_grid(q)
    pick = np.r_[0:5, 5:3125:104]  # the first five, then a stride coprime to 5 across all dimensions
    names = ("theta", "delta", "kappa_min", "kappa_step", "init_policy_random")
    lanes = {k: np.ascontiguousarray(grid[k][pick]) for k in names}
    seeds = [100 + i for i in range(len(pick))]
    op = po.OracleProblem(inst)
    pb = capi.Problem(inst)
    bt = pb.batch(len(pick), q)
    bt.set_lane_params(**lanes)
    bt.begin(seeds)
    out = bt.solve()
    for l, s in enumerate(seeds):
        over = dict(base)
        over.update({k: float(lanes[k][l]) for k in names})
        o = op.solve(po.clamp_params(util.params(**over)), s)
        assert out["status"][l] == o["status"], l
        assert out["remaining"][l] == o["remaining"], l
        assert out["loop"][l] == o["loop"], l
        assert out["value"][l] == o["value"], l
        assert np.array_equal(bt.best(l), o["x"]), l
    bt.close(), pb.close(), op.close()
