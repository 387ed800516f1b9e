"""CPU-side checks of the C ABI: the library loads, exports every symbol that
include/bxg.h declares, its host-only entry points agree with the oracle, and
compute entry points fail loudly without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from baryonyx_b200 import capi, synth
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "bxg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bxg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/bxg.h but not exported"
    assert sorted(capi.EXPORTS) == names
    assert lib.bxg_version() == 2


def test_default_and_clamp_params_match_oracle():
    a, b = capi.default_params(), po.default_params()
    assert a.as_dict() == b.as_dict()
    rng = np.random.RandomState(0)
    for t in range(200):
        p = po.default_params()
        p.theta = rng.choice([-0.1, 0.0, 0.3, 1.0, 1.5])
        p.delta = rng.choice([-1.0, -0.5, 0.0, 0.01])
        p.kappa_min = rng.choice([0.0, 0.1, 0.7, 1.0])
        p.kappa_step = rng.choice([-0.1, 0.0, 1e-3, 1.0])
        p.kappa_max = rng.choice([0.05, 0.6, 1.0, 1.2])
        p.alpha = rng.choice([-1.0, 0.0, 1.0, 2.0])
        p.w = rng.choice([-3.0, 0.0, 0.05, 0.999, 1.0, 20.0, 5000.0])
        p.limit = int(rng.choice([-1, 0, 10, 1000]))
        p.time_limit = rng.choice([-1.0, 0.0, 10.0])
        p.pushes_limit = int(rng.choice([-5, 0, 10]))
        p.pushing_iteration_limit = int(rng.choice([-5, 0, 10]))
        p.init_population_size = int(rng.choice([1, 5, 100]))
        p.init_policy_random = rng.choice([-0.5, 0.0, 0.5, 1.0, 2.0])
        p.init_kappa_improve_start = rng.choice([-1.0, 0.1, 0.5])
        p.init_kappa_improve_stop = rng.choice([0.0, 0.2, 2.0])
        p.pushing_k_factor = rng.choice([0.9, np.inf, np.nan])
        p.init_mutation_value_mean = rng.choice([-1.0, 0.5, 3.0, np.nan])
        q = capi.Params()
        C.memmove(C.byref(q), C.byref(p), C.sizeof(q))
        ca, cb = capi.clamp_params(q).as_dict(), po.clamp_params(p).as_dict()
        for k in ca:
            va, vb = ca[k], cb[k]
            assert va == vb or (isinstance(va, float) and np.isnan(va) and np.isnan(vb)), k


def test_which_problem_type():
    lib = capi.lib()
    f = lambda a: np.asarray(a, np.int32)
    assert lib.bxg_which_problem_type(3, f([1, 1, 1]), 0) == 0
    assert lib.bxg_which_problem_type(3, f([1, -1, 1]), 0) == 1
    assert lib.bxg_which_problem_type(3, f([1, -1, 3]), 0) == 2
    assert lib.bxg_which_problem_type(3, f([1, 1, 1]), 1) == 3
    assert lib.bxg_which_problem_type(3, f([-1, 1, 1]), 1) == 4
    assert lib.bxg_which_problem_type(3, f([1, -2, 1]), 1) == 5


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail, not compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    inst = synth.small_mixed(m=6, n=12, seed=1)
    with pytest.raises(capi.BxgError) as e:
        capi.Problem(inst)
    assert e.value.code == -2  # BXG_E_CUDA
    assert "no CPU path" in str(e.value)


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under baryonyx_b200/ may use it."""
    pkg = os.path.join(ROOT, "baryonyx_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "bx_oracle" not in text and "liboracle" not in text, f
                assert "libbxref" not in text, f


def test_lp_entry_points_never_throw_or_hang(tmp_path):
    """Malformed input comes back as a status: a directory instead of a
    file, truncated sections (the reference's own loops would spin on them), a
    constant term that registers an empty name."""
    with pytest.raises(capi.BxgError):
        capi.read_lp(path=str(tmp_path))  # a directory
    for text in ("minimize x + y st x + 3 <= 5 binary x",
                 "minimize x + y st x + y <= 5 general x",
                 "minimize x + y st x + y <= 5 bounds x <= 4",
                 "minimize x + y",
                 "minimize"):
        with pytest.raises(capi.BxgError):
            capi.read_lp(text=text)


def test_problem_create_validates_the_csr_before_using_it():
    """Decreasing / non-zero-based row_ptr, column out of range, zero factor:
    BXG_E_INVALID before anything is sized from the arrays (no CUDA needed to get
    that far or not, the answer is an error either way)."""
    lib = capi.lib()
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    ok = dict(row_ptr=[0, 2, 4], col=[0, 1, 1, 2], factor=[1, 1, 1, 1], cmin=[1, 1], cmax=[1, 1])
    bad = [dict(ok, row_ptr=[1, 2, 4]), dict(ok, row_ptr=[0, 3, 2]), dict(ok, row_ptr=[0, 0, 4]),
           dict(ok, col=[0, 1, 1, 7]), dict(ok, factor=[1, 0, 1, 1])]
    for case in bad:
        h = C.c_void_p()
        rc = lib.bxg_problem_create(2, 3, i32(case["row_ptr"]), i32(case["col"]), i32(case["factor"]),
                                    i32(case["cmin"]), i32(case["cmax"]), 0, 0, 0, i32([0]),
                                    np.zeros(1, np.float64), 0.0, 0, C.byref(h))
        assert rc in (-1, -2), (case, rc)  # INVALID (or no CUDA device here)
        assert not h.value
