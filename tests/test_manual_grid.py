"""bxg_manual_grid (the manual auto-tuner's parameter grid as
lanes): the product's host-side generator against the oracle's restatement of
manual_course (lib/src/manual-optimizer.cpp:31-160), against the compiled
reference's own manual_course where oracle/_ref is built, and against a
committed golden vector (tests/golden/manual_grid.json, written by
tests/golden/make_golden_manual_grid.py from the compiled reference)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from baryonyx_b200 import capi
from oracle import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ("theta", "delta", "kappa_min", "kappa_step", "init_policy_random")

BASES = [
    {},  # the reference's defaults
    {"theta": 0.7, "delta": 0.02, "kappa_min": 0.05, "kappa_step": 2e-3, "init_policy_random": 0.3},
    {"theta": 0.0, "delta": -1.0, "kappa_min": 0.0, "kappa_step": 1e-9, "init_policy_random": 1.0},
]


def params_pair(over):
    p, q = po.default_params(), capi.default_params()
    for k, v in over.items():
        setattr(p, k, v)
        setattr(q, k, v)
    return p, q


def same(a, b):
    for k in NAMES + ("tables",):
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.shape == y.shape, k
        assert np.array_equal(x.view(np.uint64), y.view(np.uint64)), k  # bit for bit


@pytest.mark.parametrize("over", BASES)
def test_grid_equals_oracle(over):
    p, q = params_pair(over)
    g, o = capi.manual_grid(q), po.manual_grid(p)
    assert len(g["theta"]) == 5 ** 5
    same(g, o)
    # entry 0 of every table is the caller's parameter, the walk starts there
    for d, k in enumerate(NAMES):
        assert g["tables"][d, 0] == getattr(p, k)
        assert g[k][0] == getattr(p, k)
    # the last dimension runs fastest
    assert np.array_equal(g["init_policy_random"][:5], g["tables"][4])
    assert np.array_equal(g["theta"][::625], g["tables"][0])


@pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("over", BASES)
def test_grid_equals_compiled_reference(over):
    p, q = params_pair(over)
    r = po.ref_manual_grid(p)
    assert len(r["theta"]) == 5 ** 5
    same(capi.manual_grid(q), r)
    same(po.manual_grid(p), r)


def test_grid_equals_golden():
    with open(os.path.join(HERE, "golden", "manual_grid.json")) as f:
        gold = json.load(f)
    for case in gold["cases"]:
        p, q = params_pair(case["base"])
        for g in (capi.manual_grid(q), po.manual_grid(p)):
            assert [float.hex(float(v)) for v in g["tables"].reshape(-1)] == case["tables_hex"]
            idx = case["points"]
            for k in NAMES:
                assert [float.hex(float(g[k][i])) for i in idx] == case[k + "_hex"], k


def test_grid_in_slices_and_arguments():
    _, q = params_pair(BASES[1])
    whole = capi.manual_grid(q)
    parts = [capi.manual_grid(q, first, 1024) for first in range(0, 4096, 1024)]
    assert [len(s["theta"]) for s in parts] == [1024, 1024, 1024, 53]
    for k in NAMES:
        assert np.array_equal(np.concatenate([s[k] for s in parts]), whole[k])
    assert len(capi.manual_grid(q, 3125, 8)["theta"]) == 0
    lib = capi.lib()
    f = lib.bxg_manual_grid
    count = C.c_int64()
    # only some arrays wanted
    th = np.zeros(10)
    assert f(C.byref(q), 0, 10, th.ctypes.data_as(C.c_void_p), None, None, None, None, C.byref(count), None) == 0
    assert count.value == 10 and np.array_equal(th, whole["theta"][:10])
    # bad arguments are refused, with a message
    assert f(None, 0, 10, None, None, None, None, None, C.byref(count), None) != 0
    assert b"bxg_manual_grid" in lib.bxg_last_error()
    assert f(C.byref(q), -1, 10, None, None, None, None, None, C.byref(count), None) != 0
    assert f(C.byref(q), 0, 10, None, None, None, None, None, None, None) != 0
