"""The engine's preprocessor (bxg_lp_preprocess / bxg_lp_affect / bxg_lp_split,
include/bxg.h) against the reference's baryonyx::preprocess,
affect and split (lib/src/preprocessor.cpp:857-903): the same remaining
variables, rows, right-hand sides, objective, constant and fixed variables.
Host code: everything but the last test runs without a GPU."""
import ctypes as C
import glob
import json
import os

import numpy as np
import pytest

from baryonyx_b200 import capi
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PP_DIR = os.path.join(ROOT, "tests", "golden", "lp_pp")
REF_DIR = "/root/reference"
ARRAYS = ("row_ptr", "col", "factor", "cmin", "cmax", "obj_var", "obj_factor")
SCALARS = ("m", "n", "problem_type", "maximize", "cost_constant")
GOLDEN = sorted(glob.glob(os.path.join(PP_DIR, "*.lp")))
needs_ref = pytest.mark.skipif(not (os.path.isdir(REF_DIR) and po.ref_available()),
                               reason="needs /root/reference and oracle/_ref (build container)")


def same_problem(mine: dict, want: dict, what):
    for k in SCALARS:
        assert mine[k] == want[k], (what, k, mine[k], want[k])
    for k in ARRAYS:
        a, b = np.asarray(mine[k]), np.asarray(want[k])
        assert a.shape == b.shape and np.array_equal(a, b), (what, k)
    assert mine["names"] == list(want["names"]), (what, "names")
    assert [tuple(b) for b in mine["bounds"]] == [tuple(b) for b in want["bounds"]], (what, "bounds")
    assert [tuple(a) for a in mine["affected"]] == [tuple(a) for a in want["affected"]], (what, "affected")


def engine(lp, pre_order, affect=None, text=None):
    """("ok", problem) or ("refuses", message)"""
    try:
        return "ok", capi.read_lp(path=lp, text=text, pre_order=pre_order, preprocess=True, affect=affect)
    except capi.BxgError as e:
        assert e.code == -1  # BXG_E_INVALID
        return "refuses", str(e)


def test_golden_fixtures_present():
    assert len(GOLDEN) >= 4
    for lp in GOLDEN:
        assert os.path.exists(lp[:-3] + ".json"), lp


@pytest.mark.parametrize("lp", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_preprocess_and_affect_match_reference_golden(lp):
    doc = json.load(open(lp[:-3] + ".json"))
    assert doc
    for key, want in doc.items():
        parts = key.split("_")
        pre_order = int(parts[2])
        affect = [(int(parts[4]), int(parts[5]))] if len(parts) > 3 else None
        status, mine = engine(lp, pre_order, affect)
        if want["status"] == "ok":
            assert status == "ok", (key, mine)
            same_problem(mine, want["problem"], (lp, key))
        else:
            # bx_reach() / a failed contract in the reference: an error return here
            assert status == "refuses", key
            assert "preprocessor" in mine


def test_every_rule_fires_on_the_chain_fixture():
    """What each rule of the reference does to tests/golden/lp_pp/chain_fixing.lp
    (preprocessor.cpp:441-559, :337-358), spelled out."""
    d = capi.read_lp(path=os.path.join(PP_DIR, "chain_fixing.lp"), preprocess=True)
    fixed = dict(d["affected"])
    assert fixed["s"] == 0                                  # bound s = 0
    assert fixed["p"] == 1                                  # one-element row
    assert fixed["q"] == 0                                  # p + q <= 1 counted down to one
    assert fixed["r"] == 1                                  # q + r + s >= 1 with q = s = 0
    assert (fixed["u0"], fixed["u1"], fixed["u2"]) == (1, 1, 1)   # sum of factors == value
    assert (fixed["k1"], fixed["k2"], fixed["k3"]) == (0, 0, 0)   # <= 0 over positive factors
    assert (fixed["g1"], fixed["g2"]) == (1, 1)
    assert (fixed["h1"], fixed["h2"], fixed["h3"]) == (0, 0, 0)   # row always true: unused, positive costs
    assert fixed["free_neg"] == 0 and fixed["free_pos"] == 1      # unused: the value the cost likes
    assert fixed["nocost"] == 1                                   # unused and not in the objective: true
    assert d["names"] == ["v1", "v2", "v3"]
    # keep3: v2 - v3 + r <= 1 with r = 1
    assert d["m"] == 3 and d["cmax"].tolist()[-1] == 0
    # objective constant: -6.5 + the costs of the variables fixed to one
    ones = {"p": 3, "r": -4, "u0": 7, "u1": 1, "u2": -2, "g1": 11, "g2": 12, "free_pos": -3}
    assert d["cost_constant"] == -6.5 + sum(ones.values())


def test_split_is_the_two_affects():
    lp = os.path.join(PP_DIR, "maximize_general.lp")
    a, b = capi.split_lp(lp, 0)
    same_problem(a, capi.read_lp(path=lp, preprocess=True, affect=[(0, 1)]), "split true")
    same_problem(b, capi.read_lp(path=lp, preprocess=True, affect=[(0, 0)]), "split false")
    assert len(a["affected"]) > len(capi.read_lp(path=lp, preprocess=True)["affected"])


def test_affect_refuses_an_index_out_of_range():
    """preprocessor.cpp:866,884: bx_expects on the variable index"""
    lp = os.path.join(PP_DIR, "maximize_general.lp")
    n = capi.read_lp(path=lp, preprocess=True)["n"]
    for v in (-1, n, n + 7):
        with pytest.raises(capi.BxgError):
            capi.read_lp(path=lp, preprocess=True, affect=[(v, 1)])


def test_bound_affectation():
    """lib/test/preprocess.cpp:163-204 (test_bound_affectation): nothing to remove
    until a bound fixes one variable; then that variable alone goes."""
    text = ("minimize\n a + b + c + d + e + f\nst\n c1: a + b + c <= 2\n c2: d + e + f <= 2\n"
            "%s\nend\n")
    d0 = capi.read_lp(text=text % "binary\n a b c d e f", preprocess=True)
    assert d0["n"] == 6 and d0["affected"] == [] and d0["m"] == 2
    # (a binary section after the bounds would reset f to [0, 1], as in the reference)
    d1 = capi.read_lp(text=text % "bounds\n f = 0\nbinary\n a b c d e", preprocess=True)
    assert d1["n"] == 5 and d1["affected"] == [("f", 0)] and d1["m"] == 2
    assert d1["cmax"].tolist() == [2, 2]


@needs_ref
def test_reference_own_preprocess_cases():
    """lib/test/preprocess.cpp:206-300 on the reference's prepro.lp: 23 variables
    -> 2 remaining and 21 fixed; affect / split of either remaining variable
    fixes the other one as well."""
    lp = os.path.join(REF_DIR, "lib", "test", "prepro.lp")
    d = capi.read_lp(path=lp, preprocess=True)
    assert d["n"] == 2 and len(d["affected"]) == 21
    for v in (0, 1):
        for b in (1, 0):
            a = capi.read_lp(path=lp, preprocess=True, affect=[(v, b)])
            assert a["n"] == 0 and len(a["affected"]) == 23
    p0, p1 = capi.split_lp(lp, 0)
    assert p0["n"] == 0 and len(p0["affected"]) == 23
    assert p1["n"] == 0 and len(p1["affected"]) == 23


REF_LPS = sorted(glob.glob(os.path.join(REF_DIR, "**", "*.lp"), recursive=True))


@needs_ref
def test_every_reference_fixture_preprocesses_like_the_reference():
    """All LP files the reference ships through baryonyx::preprocess of the
    compiled reference (and one affect on the result), four pre_constraint_order
    values; where the reference's preprocessor terminates (vm.lp: bx_reach) the
    engine returns an error."""
    assert len(REF_LPS) >= 50
    same = refused = 0
    for lp in REF_LPS:
        big = os.path.getsize(lp) > 400_000
        for pre_order in ((0,) if big else (0, 1, 5, 9)):
            p = po.default_params()
            p.pre_order = pre_order
            for affect in ((None,) if big else (None, [(0, 1)], [(0, 0)])):
                status, want = po.ref_preprocess(lp, p, affect)
                mine_status, mine = engine(lp, pre_order, affect)
                if status == "ok":
                    assert mine_status == "ok", (lp, pre_order, affect, mine)
                    same_problem(mine, want, (lp, pre_order, affect))
                    same += 1
                else:
                    assert mine_status == "refuses", (lp, pre_order, affect, status)
                    refused += 1
    assert same >= 300 and refused >= 1


def _random_lp(rng) -> str:
    """Small binary problems dense in what the preprocessor looks for: fixing
    bounds, one-element rows, rows whose factors sum to the right-hand side,
    zero right-hand sides, short chains, negative factors."""
    n = int(rng.integers(3, 12))
    names = [f"x{i}" for i in range(n)]
    lines = ["maximize" if rng.random() < 0.4 else "minimize"]
    obj = [f"{'+' if rng.random() < 0.6 else '-'} {int(rng.integers(1, 9))} {v}" for v in names if rng.random() < 0.8]
    lines.append(" " + " ".join(obj or ["x0"]))
    lines.append("st")
    for r in range(int(rng.integers(1, 9))):
        k = int(rng.integers(1, min(n, 4) + 1))
        picks = rng.choice(n, size=k, replace=False)
        positive = rng.random() < 0.75
        facs = [int(rng.integers(1, 3)) if positive or rng.random() < 0.5 else -int(rng.integers(1, 3)) for _ in picks]
        total = sum(f for f in facs if f > 0)
        rhs = int(rng.choice([0, 1, total, max(total - 1, 0), int(rng.integers(0, 3))]))
        op = rng.choice(["=", "<=", ">="])
        lhs = " ".join(f"{'+' if f > 0 else '-'} {abs(f) if abs(f) != 1 else ''} {names[int(j)]}" for f, j in zip(facs, picks))
        lines.append(f" c{r}: {lhs} {op} {rhs}")
    fixed = [v for v in names if rng.random() < 0.15]
    if fixed:
        lines.append("bounds")
        lines += [f" {v} = {int(rng.integers(0, 2))}" for v in fixed]
    lines += ["binary", " " + " ".join(names), "end", ""]
    return "\n".join(lines)


@needs_ref
def test_generated_problems_preprocess_like_the_reference(tmp_path):
    """400 generated texts: the compiled reference's preprocess (+ one affect)
    and the engine's agree on the shrunk problem, or both refuse."""
    rng = np.random.default_rng(20261018)
    same = refused = 0
    for i in range(400):
        text = _random_lp(rng)
        f = tmp_path / f"g{i}.lp"
        f.write_text(text)
        pre_order = int(rng.integers(0, 12))
        p = po.default_params()
        p.pre_order = pre_order
        status, want = po.ref_preprocess(str(f), p)
        mine_status, mine = engine(str(f), pre_order)
        if status != "ok":
            assert mine_status == "refuses", (text, status, mine)
            refused += 1
            continue
        assert mine_status == "ok", (text, mine)
        same_problem(mine, want, text)
        same += 1
        if want["n"] > 0:
            affect = [(int(rng.integers(0, want["n"])), int(rng.integers(0, 2)))]
            status, want = po.ref_preprocess(str(f), p, affect)
            mine_status, mine = engine(str(f), pre_order, affect)
            if status == "ok":
                assert mine_status == "ok", (text, affect, mine)
                same_problem(mine, want, (text, affect))
            else:
                assert mine_status == "refuses", (text, affect, status)
    assert same >= 150, (same, refused)


def _satisfies(raw: dict, value_of: dict) -> bool:
    x = np.array([value_of[name] for name in raw["names"]], np.int64)
    for k in range(raw["m"]):
        a, b = raw["row_ptr"][k], raw["row_ptr"][k + 1]
        v = int(np.dot(raw["factor"][a:b].astype(np.int64), x[raw["col"][a:b]]))
        if not (raw["cmin"][k] <= v <= raw["cmax"][k]):
            return False
    return True


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["chain_fixing", "maximize_general"])
def test_preprocessed_problem_solves_and_lifts_to_the_original(name):
    """LP text -> bxg_lp_preprocess -> bxg_problem_create -> solve on the device:
    equal to the oracle on the golden (reference-made) shrunk problem, and the
    solution together with the fixed variables satisfies every row of the
    ORIGINAL problem, with the original objective value."""
    lp = os.path.join(PP_DIR, name + ".lp")
    raw = capi.read_lp(path=lp)
    inst = capi.read_lp(path=lp, preprocess=True)
    want = json.load(open(lp[:-3] + ".json"))["pre_order_0"]["problem"]
    oinst = {k: (np.asarray(v, np.float64 if k == "obj_factor" else np.int32) if k in ARRAYS else v)
             for k, v in want.items()}
    p = po.default_params()
    p.limit = 300
    q = capi.Params()
    C.memmove(C.byref(q), C.byref(p), C.sizeof(q))
    seeds = [5, 99]
    op = po.OracleProblem(oinst)
    pb = capi.Problem(inst)
    bt = pb.batch(len(seeds), q)
    bt.begin(seeds)
    out = bt.solve()
    solved = 0
    for lane, seed in enumerate(seeds):
        sol = op.solve(p, seed)
        assert int(out["status"][lane]) == sol["status"], (name, lane)
        if sol["status"] != 0:
            continue
        x = bt.best(lane)
        assert np.array_equal(x, sol["x"]), (name, lane)
        assert float(out["value"][lane]) == sol["value"], (name, lane)
        value_of = dict(inst["affected"])
        value_of.update({n_: int(b) for n_, b in zip(inst["names"], x)})
        assert _satisfies(raw, value_of), (name, lane)
        cost = raw["cost_constant"] + sum(f * value_of[raw["names"][v]] for v, f in zip(raw["obj_var"], raw["obj_factor"]))
        assert abs(cost - float(out["value"][lane])) < 1e-9, (name, lane, cost)
        solved += 1
    assert solved >= 1
    bt.close()
    op.close()
    pb.close()
