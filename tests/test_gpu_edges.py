"""Edge cases of the sweep engine against the oracle: the smallest shapes,
bit_array block boundaries (n = 63, 64, 65, 128), rows of one element, a
variable in no constraint, zero costs, more lanes than one wave of CTAs."""
import numpy as np
import pytest

from baryonyx_b200 import capi, synth
from oracle import pyoracle as po
from tests import util
from tests.test_gpu_parity import lockstep_engine, solve_case, to_capi

pytestmark = pytest.mark.gpu

INT_MIN, INT_MAX = -(2**31), 2**31 - 1


def inst_of(n, rows, factors, cmin, cmax, costs, maximize=False):
    return synth._finish(len(rows), n, rows, factors, cmin, cmax, costs, maximize=maximize, name="edge")


def test_single_row_and_single_element_rows():
    # one equality row over three variables
    inst = inst_of(3, [[0, 1, 2]], [[1, 1, 1]], [1], [1], [3, 1, 2])
    lockstep_engine(inst, util.params(), [1, 2], 6, pushes_every=3)
    solve_case(inst, util.params(limit=20, pushes_limit=2, pushing_iteration_limit=2), [1, 2, 3])
    # rows of a single element next to longer ones, one variable in no row
    rows = [[0], [1, 2], [3], [0, 2, 4, 5], [5]]
    fac = [[1], [1, 1], [1], [1, 1, 1, 1], [1]]
    inst = inst_of(7, rows, fac, [1, 1, 0, 1, INT_MIN], [1, INT_MAX, 0, 2, 1], [1, 1, 2, 2, 0, 3, 5])
    lockstep_engine(inst, util.params(kappa_step=0.01), [4, 5, 6], 12, pushes_every=4)
    solve_case(inst, util.params(limit=40, pushes_limit=3, pushing_iteration_limit=3), [4, 5, 6])


@pytest.mark.parametrize("n", [63, 64, 65, 127, 128, 129])
def test_bit_array_block_boundaries(n):
    # every variable sits in two rows; the last variable closes a 64-bit block or opens one
    m = 12
    rng = np.random.RandomState(n)
    rows = [[] for _ in range(m)]
    for j in range(n):
        for k in rng.choice(m, size=2, replace=False):
            rows[k].append(j)
    rows = [sorted(r) for r in rows]
    fac = [[1] * len(r) for r in rows]
    inst = inst_of(n, rows, fac, [1] * m, [INT_MAX] * m, rng.randint(0, 4, size=n))  # zero costs included
    lockstep_engine(inst, util.params(kappa_step=0.01), [7, 8], 10, pushes_every=5)
    inst = inst_of(n, rows, fac, [2] * m, [2] * m, rng.randint(1, 3, size=n), maximize=True)
    lockstep_engine(inst, util.params(kappa_step=0.01), [7, 8], 10, pushes_every=5)


def test_more_lanes_than_one_wave():
    """1500 lanes of a small instance (more CTAs than fit the GPU at once):
    a sample of lanes against the oracle, all lanes against their seed twins."""
    inst = synth.spp(m=40, n=300, lo=3, hi=6, seed=2, cost_hi=5)
    cp = po.clamp_params(util.params(limit=40, pushes_limit=2, pushing_iteration_limit=3))
    L = 1500
    seeds = (np.arange(L, dtype=np.uint32) % 750) + 5  # lane l and l + 750 share a seed
    pb = capi.Problem(inst)
    bt = pb.batch(L, to_capi(cp))
    bt.begin(seeds)
    out = bt.solve()
    for k in ("status", "remaining", "loop"):
        assert np.array_equal(out[k][:750], out[k][750:]), k
    op = po.OracleProblem(inst)
    for l in (0, 1, 147, 148, 749, 750, 1036, 1499):
        o = op.solve(cp, int(seeds[l]))
        assert out["status"][l] == o["status"] and out["loop"][l] == o["loop"], l
        assert np.array_equal(bt.best(l), o["x"]), l
    bt.close(), pb.close(), op.close()


@pytest.mark.parametrize("m, n, lo, hi, sweeps", [
    (8, 20000, 2, 4, 8),    # rows of ~7.5 k elements: two chunks of the shared-memory buffer
    (6, 60000, 3, 5, 5),    # rows of ~40 k: three chunks, strides that cross chunks in global memory
    (4, 64000, 3, 4, 4),    # rows of ~56 k: close to the longest the CTA-wide sort takes (16-bit positions)
])
def test_rows_whose_scratch_is_in_global_memory(m, n, lo, hi, sweeps):
    """Rows too long for a shared-memory scratch (update_row's direct gather,
    cta_bitonic_chunked, the tie-table shortcut): integer costs, so equal values
    everywhere, +-1 factors; state equal to the oracle's after every sweep."""
    inst = synth.csplib022(m=m, n=n, lo=lo, hi=hi, seed=31 + m, name="long-rows")
    lockstep_engine(inst, util.params(kappa_step=0.01), [3, 4], sweeps, pushes_every=3)
    # inequality rows of the same shape (selection window, stopping values)
    inst = synth.csplib022(m=m, n=n, lo=lo, hi=hi, seed=41 + m, name="long-rows-ineq")
    inst["cmin"] = (np.asarray(inst["cmin"]) - 1).astype(np.int32)
    inst["problem_type"] = 4  # inequalities_101
    lockstep_engine(inst, util.params(kappa_step=0.02), [5], sweeps)


@pytest.mark.parametrize("decimals", [None, 3, 1])
def test_long_rows_rank_select(decimals):
    """Long rows whose values differ, or with a few equal values only (real
    costs, rounded to `decimals` places): the boundary inside the row is found
    by the radix select (update_row mode 4, cta_rank_select) and the equal
    values listed by the tie table draw from the engine in ascending order;
    state equal to the oracle's after every sweep, equality and inequality rows."""
    rng = np.random.RandomState(77)
    for ineq in (False, True):
        inst = synth.csplib022(m=8, n=20000, lo=2, hi=4, seed=51, name="long-rows-real-costs")
        c = rng.rand(inst["n"]) * 50 + 1
        if decimals is not None:
            c = np.round(c, decimals)
        inst["obj_factor"] = c.astype(np.float64)
        if ineq:
            inst["cmin"] = (np.asarray(inst["cmin"]) - 1).astype(np.int32)
            inst["problem_type"] = 4  # inequalities_101
        lockstep_engine(inst, util.params(kappa_step=0.01), [3, 4], 8, pushes_every=3)
