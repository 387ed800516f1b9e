"""The engine's LP reader (bxg_lp_*, include/bxg.h) against the reference's
make_problem + make_merged_constraints (lib/src/parser.cpp:1065-1273,
lib/src/merged-constraint.cpp:297-343): same variable numbering, rows, bounds,
objective and solver family.  Host code: these run without a GPU; the last
test feeds a parsed file to the CUDA engine (-m gpu)."""
import ctypes as C
import glob
import json
import os

import numpy as np
import pytest

from baryonyx_b200 import capi, synth
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LP_DIR = os.path.join(ROOT, "tests", "golden", "lp")
REF_DIR = "/root/reference"
INT_MIN, INT_MAX = -(2**31), 2**31 - 1
ARRAYS = ("row_ptr", "col", "factor", "cmin", "cmax", "obj_var", "obj_factor")
SCALARS = ("m", "n", "problem_type", "maximize", "cost_constant")


def same_parse(mine: dict, want: dict, what):
    for k in SCALARS:
        assert mine[k] == want[k], (what, k, mine[k], want[k])
    for k in ARRAYS:
        a, b = np.asarray(mine[k]), np.asarray(want[k])
        assert a.shape == b.shape and np.array_equal(a, b), (what, k)
    if "names" in want:
        assert mine["names"] == list(want["names"]), (what, "names")
        assert [tuple(b) for b in mine["bounds"]] == [tuple(b) for b in want["bounds"]], (what, "bounds")


GOLDEN = sorted(glob.glob(os.path.join(LP_DIR, "*.lp")))


def test_golden_lp_fixtures_present():
    assert len(GOLDEN) >= 4
    for lp in GOLDEN:
        assert os.path.exists(lp[:-3] + ".json"), lp


@pytest.mark.parametrize("lp", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
@pytest.mark.parametrize("pre_order", list(range(12)))
def test_reader_matches_reference_golden(lp, pre_order):
    want = json.load(open(lp[:-3] + ".json"))[f"pre_order_{pre_order}"]
    same_parse(capi.read_lp(path=lp, pre_order=pre_order), want, lp)
    # same text through the in-memory entry point
    same_parse(capi.read_lp(text=open(lp).read(), pre_order=pre_order), want, lp)


def test_named_objective():
    """lib/test/io.cpp:42-61"""
    d = capi.read_lp(text="maximize\nx0: +x1 + 2x2 + 3x3 - 100\nend\n")
    assert d["maximize"] == 1 and d["m"] == 0 and d["n"] == 3
    assert d["obj_var"].tolist() == [0, 1, 2] and d["obj_factor"].tolist() == [1.0, 2.0, 3.0]
    assert d["cost_constant"] == -100.0
    assert d["names"] == ["x1", "x2", "x3"]


def test_unnamed_objective_goes_to_constraints():
    """lib/test/io.cpp:63-80"""
    d = capi.read_lp(text="maximize\nst: x1 + x2 + x3 = 1\nend\n")
    assert d["maximize"] == 1 and len(d["obj_var"]) == 0
    assert d["n"] == 3 and d["m"] == 1
    assert d["col"].tolist() == [0, 1, 2] and d["factor"].tolist() == [1, 1, 1]
    assert d["cmin"].tolist() == [1] and d["cmax"].tolist() == [1]


def test_small_lp():
    """lib/test/io.cpp:82-150; the two rows over x1 - 3x2 + x3 merge into one
    (merged-constraint.cpp:52-175: the tighter bound wins)."""
    text = ("maximize\nobj: x1 + 2x2 + 3x3 - 100\nst\ntime:  -x1 + x2 + x3 <= 20\n"
            "labor:  x1 - 3x2 + x3 <= 30\ntest: x1 - 3x2 + x3 <= -5\nbounds\nx1 <= 40\nend\n")
    d = capi.read_lp(text=text)
    assert d["obj_factor"].tolist() == [1.0, 2.0, 3.0] and d["cost_constant"] == -100.0
    assert d["names"] == ["x1", "x2", "x3"]
    assert [b[0] for b in d["bounds"]] == [0, 0, 0]
    assert [b[1] for b in d["bounds"]] == [40, INT_MAX, INT_MAX]
    assert d["m"] == 2
    assert d["factor"].tolist() == [-1, 1, 1, 1, -3, 1]
    assert d["cmin"].tolist() == [INT_MIN, INT_MIN] and d["cmax"].tolist() == [20, -5]


@pytest.mark.parametrize("text, tag, reference", [
    ("", "bad_objective_function_type", "refuses"),
    ("minimise\n x + y\nend\n", "bad_objective_function_type", "refuses"),
    # no right-hand side: the reference's pop_front(0) contract calls std::terminate
    # (parser.cpp:302); behind a C ABI that is an error return
    ("minimize\n x + y\nst\n c: x + y <= \nend\n", "bad_constraint", "aborts"),
    ("minimize\n x + y\nst\n c: x + y <= 1\nbounds\n x <= \nend\n", "bad_bound", "refuses"),
    # a variable called e reads as a number in the bounds section (parser.cpp read_real / sscanf)
    ("minimize\n a + e\nst\n c: a + e <= 1\nbounds\n e <= 1\nend\n", "bad_bound", "refuses"),
    # the general section comes after the binary section in this dialect
    ("minimize\n a + b\nst\n c: a + b <= 1\ngeneral\n a\nbinary\n b\nend\n", "bad_end", "refuses"),
    ("minimize\n x + y\nst\n c: x + y <= 1\n", "bad_end", "refuses"),
    # quadratic objectives parse in the reference; no solver of this path takes them
    ("minimize\n x + [ x * y ] / 2\nst\n c: x + y <= 1\nend\n", "quadratic", "accepts"),
])
def test_malformed_text_is_refused_with_the_reference_tag(text, tag, reference, tmp_path):
    with pytest.raises(capi.BxgError) as e:
        capi.read_lp(text=text)
    assert tag in str(e.value), str(e.value)
    assert e.value.code == -1  # BXG_E_INVALID
    if po.ref_available() and reference == "refuses":
        f = tmp_path / "bad.lp"
        f.write_text(text)
        with pytest.raises(RuntimeError):
            po.RefProblem(str(f), po.default_params())


def test_per_variable_and_bulk_queries_agree():
    """bxg_lp_variable (one at a time) against bxg_lp_variables / bxg_lp_names."""
    lp = os.path.join(LP_DIR, "general_bounds.lp")
    d = capi.read_lp(path=lp)  # bulk
    L = capi.lib()
    h = C.c_void_p()
    capi.check(L.bxg_lp_read(lp.encode(), 0, C.byref(h)))
    try:
        for i in range(d["n"]):
            lo, hi, ty = C.c_int(), C.c_int(), C.c_int()
            name = L.bxg_lp_variable(h, i, C.byref(lo), C.byref(hi), C.byref(ty))
            assert name.decode() == d["names"][i]
            assert (lo.value, hi.value, ty.value) == d["bounds"][i]
        assert L.bxg_lp_variable(h, d["n"], None, None, None) is None
        need = C.c_size_t()
        small = C.create_string_buffer(2)
        assert L.bxg_lp_names(h, small, 2, C.byref(need)) != 0  # too small: refused, nothing written
        assert need.value == sum(len(n) + 1 for n in d["names"])
    finally:
        L.bxg_lp_destroy(h)


def test_missing_file():
    with pytest.raises(capi.BxgError) as e:
        capi.read_lp(path="/nonexistent/file.lp")
    assert "file_not_found" in str(e.value)


def test_unknown_pre_order():
    for bad in (-2, 12):
        with pytest.raises(capi.BxgError):
            capi.read_lp(text="minimize\n x\nst\n c: x <= 1\nend\n", pre_order=bad)


def test_raw_parse_reports_no_rows():
    """BXG_LP_RAW (-1): the parse alone, the input of bxg_lp_preprocess; rows
    that contradict each other are not an error before the merge."""
    text = "minimize\n x + y\nst\n a: x + y = 1\n b: x + y = 2\nbinary\n x y\nend\n"
    d = capi.read_lp(text=text, pre_order=-1)
    assert d["m"] == 0 and d["n"] == 2 and d["names"] == ["x", "y"] and d["problem_type"] == 0
    with pytest.raises(capi.BxgError) as e:
        capi.read_lp(text=text, pre_order=0)
    assert "multiple_constraint" in str(e.value)


@pytest.mark.parametrize("make", [
    lambda: synth.small_mixed(family="01", equality=True, seed=11),
    lambda: synth.small_mixed(family="101", seed=12),
    lambda: synth.small_mixed(family="Z", seed=13, maximize=True),
    lambda: synth.scp(),
    lambda: synth.spp(m=100, n=2000),
    lambda: synth.telebus(m=150, n=700),
])
def test_write_then_read_round_trip(make, tmp_path):
    """synth.write_lp -> bxg_lp_read returns the instance it was written from
    (size-independent property; BASELINE-shaped generators included)."""
    inst = make()
    path = synth.write_lp(inst, str(tmp_path / "rt.lp"))
    d = capi.read_lp(path=path)
    assert d["n"] == inst["n"] and d["maximize"] == inst["maximize"]
    assert d["problem_type"] == inst["problem_type"]
    cost = np.zeros(inst["n"])
    np.add.at(cost, np.asarray(inst["obj_var"]), np.asarray(inst["obj_factor"]))
    mine = np.zeros(inst["n"])
    np.add.at(mine, d["obj_var"], d["obj_factor"])
    assert np.array_equal(cost, mine)
    for k in ("m", "row_ptr", "col", "factor", "cmin", "cmax"):
        assert np.array_equal(np.asarray(d[k]), np.asarray(inst[k])), k
    assert all(b == (0, 1, 1) for b in d["bounds"])


REF_LPS = sorted(glob.glob(os.path.join(REF_DIR, "**", "*.lp"), recursive=True))


@pytest.mark.skipif(not (REF_LPS and po.ref_available()), reason="needs /root/reference and oracle/_ref (build container)")
def test_every_reference_fixture_parses_like_the_reference():
    """All LP files the reference ships (lib/test/*.lp and the rest), all twelve
    pre_constraint_order values, against the compiled reference's parser and
    make_merged_constraints."""
    assert len(REF_LPS) >= 50
    compared = 0
    for lp in REF_LPS:
        for pre_order in range(12):
            p = po.default_params()
            p.pre_order = pre_order
            try:
                r = po.RefProblem(lp, p)
            except RuntimeError:
                with pytest.raises(capi.BxgError):
                    capi.read_lp(path=lp, pre_order=pre_order)
                continue
            want = r.instance()
            want["names"], want["bounds"] = r.variables()
            r.close()
            same_parse(capi.read_lp(path=lp, pre_order=pre_order), want, (lp, pre_order))
            compared += 1
    assert compared >= 600


@pytest.mark.gpu
@pytest.mark.parametrize("lp", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_parsed_file_solves_like_the_oracle(lp):
    """LP text -> bxg_lp_* -> bxg_problem_create -> solve on the device equals
    the oracle on the golden parse of the same file."""
    inst = capi.read_lp(path=lp)
    want = json.load(open(lp[:-3] + ".json"))["pre_order_0"]
    oinst = {k: (np.asarray(v, np.float64 if k == "obj_factor" else np.int32) if k in ARRAYS else v)
             for k, v in want.items()}
    p = po.default_params()
    p.limit = 200
    q = capi.Params()
    C.memmove(C.byref(q), C.byref(p), C.sizeof(q))
    seeds = [3, 17, 123456]
    if (np.asarray(want["factor"]) == 0).any():
        # factors that cancel stay in the row with factor 0, as in the reference's
        # parse; the engine refuses such a row loudly (bxg_problem_create)
        with pytest.raises(capi.BxgError) as e:
            capi.Problem(inst)
        assert "zero factor" in str(e.value)
        return
    op = po.OracleProblem(oinst)
    pb = capi.Problem(inst)
    bt = pb.batch(len(seeds), q)
    bt.begin(seeds)
    out = bt.solve()
    for lane, seed in enumerate(seeds):
        sol = op.solve(p, seed)
        assert int(out["status"][lane]) == sol["status"], (lp, lane)
        if sol["status"] == 0:
            assert np.array_equal(bt.best(lane), sol["x"]), (lp, lane)
            assert float(out["value"][lane]) == sol["value"], (lp, lane)
    bt.close()
    op.close()
    pb.close()


def _random_lp(rng) -> str:
    """Well-formed LP text with the dialect's variety: labels or none, decimal
    and exponent coefficients, repeated variables, the operator spellings,
    the bound forms, binary / general sections."""
    n = int(rng.integers(2, 9))
    names = [rng.choice(["x", "y_", "var.", "Q"]) + str(int(rng.integers(0, 50))) + rng.choice(["", "a", "_z"])
             for _ in range(n)]
    names = list(dict.fromkeys(names))

    def coef(first):
        v = rng.choice([1, 1, 1, 2, 3, 2.5, 10, 0.5, 7])
        sign = rng.choice(["+", "-"])
        txt = {2.5: "2.5", 0.5: "0.5", 10: rng.choice(["10", "1e1"])}.get(v, str(int(v)))
        if v == 1 and rng.random() < 0.6:
            txt = ""
        lead = "" if (first and sign == "+" and rng.random() < 0.5) else sign + rng.choice([" ", ""])
        return lead + txt + (" " if txt and rng.random() < 0.7 else "")

    def expr(k):
        picks = [names[int(i)] for i in rng.integers(0, len(names), size=k)]
        return " ".join(coef(i == 0) + v for i, v in enumerate(picks))

    out = [rng.choice(["minimize", "maximize", "Minimize", "MAX", "min"])]
    if rng.random() < 0.8:
        label = "obj: " if rng.random() < 0.5 else ""
        const = rng.choice(["", " + 4", " - 2.5"])
        out.append(" " + label + expr(int(rng.integers(1, 6))) + const)
    out.append(rng.choice(["subject to", "st", "s.t.", "Subject To"]))
    for k in range(int(rng.integers(1, 7))):
        label = f"c{k}: " if rng.random() < 0.6 else ""
        op = rng.choice(["<=", ">=", "=", "=<", "=>", "<", ">"])
        rhs = rng.choice(["0", "1", "2", "3", "-1", "2.7", "+2"])
        out.append(" " + label + expr(int(rng.integers(1, 5))) + " " + op + " " + rhs)
    kind = rng.choice(["binary", "bounds", "general", "none"])
    if kind == "bounds":
        out.append("bounds")
        for v in names:
            form = rng.choice(["up", "both", "skip"])
            if form == "up":
                out.append(f" {v} <= {int(rng.integers(1, 5))}")
            elif form == "both":
                out.append(f" {int(rng.integers(0, 2))} <= {v} <= {int(rng.integers(2, 6))}")
    elif kind == "binary":
        out.append(rng.choice(["binary", "binaries", "bin"]))
        out.append(" " + " ".join(names))
    elif kind == "general":
        out.append("bounds")
        out.extend(f" {v} <= 3" for v in names)
        out.append(rng.choice(["general", "generals", "gen"]))
        out.append(" " + " ".join(names))
    out.append(rng.choice(["end", "End"]))
    return "\n".join(out) + "\n"


@pytest.mark.skipif(not (os.path.isdir(REF_DIR) and po.ref_available()), reason="needs oracle/_ref (build container)")
def test_random_texts_parse_like_the_reference(tmp_path):
    """Differential run: 300 generated LP texts, a random constraint order
    each, the engine's reader against the compiled reference's."""
    rng = np.random.default_rng(20261018)
    compared = refused = 0
    for case in range(300):
        text = _random_lp(rng)
        pre_order = int(rng.integers(0, 12))
        f = tmp_path / f"r{case}.lp"
        f.write_text(text)
        p = po.default_params()
        p.pre_order = pre_order
        try:
            r = po.RefProblem(str(f), p)
        except RuntimeError:
            with pytest.raises(capi.BxgError):
                capi.read_lp(text=text, pre_order=pre_order)
            refused += 1
            continue
        want = r.instance()
        want["names"], want["bounds"] = r.variables()
        r.close()
        same_parse(capi.read_lp(text=text, pre_order=pre_order), want, (case, pre_order, text))
        compared += 1
    assert compared >= 150, (compared, refused)
