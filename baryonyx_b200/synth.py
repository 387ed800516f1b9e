"""Synthetic instances of the shapes BASELINE.json names.

The real benchmark files (OR-lib scp41, SPP, csplib022, telebus, rail4284) are
not in the reference repository (benchmarks/download.sh fetches them), so the
bench and the parity tests use generated instances of the same *shape*.  Every
generator is seeded and returns the same arrays on every machine
(``numpy.random.RandomState``).

An instance is a dict of merged constraints in CSR form, i.e. what the
reference's ``make_merged_constraints`` (lib/src/merged-constraint.cpp:297-343)
hands to a Solver constructor, plus the objective in file order:

    m, n, row_ptr[m+1], col[nnz], factor[nnz], cmin[m], cmax[m],
    problem_type (baryonyx::problem_solver_type), maximize,
    obj_var[], obj_factor[], cost_constant

``write_lp`` emits the same instance as CPLEX-LP text so that the reference's
own reader can parse it (bench ``--impl reference`` and the parity tests).
"""
from __future__ import annotations

import numpy as np

INT_MIN = -(2**31)
INT_MAX = 2**31 - 1

EQ_01, EQ_101, EQ_Z, INEQ_01, INEQ_101, INEQ_Z = range(6)


def _finish(m, n, rows, factors, cmin, cmax, costs, maximize=False, name=""):
    row_ptr = np.zeros(m + 1, np.int32)
    for k in range(m):
        row_ptr[k + 1] = row_ptr[k] + len(rows[k])
    col = np.concatenate([np.asarray(r, np.int32) for r in rows]) if m else np.zeros(0, np.int32)
    fac = np.concatenate([np.asarray(f, np.int32) for f in factors]) if m else np.zeros(0, np.int32)
    cmin = np.asarray(cmin, np.int64)
    cmax = np.asarray(cmax, np.int64)
    has_ineq = bool(np.any(cmin != cmax))
    if np.any(np.abs(fac) >= 2):
        coef = 2
    elif np.any(fac == -1):
        coef = 1
    else:
        coef = 0
    ptype = (INEQ_01, INEQ_101, INEQ_Z)[coef] if has_ineq else (EQ_01, EQ_101, EQ_Z)[coef]
    return {
        "name": name,
        "m": int(m),
        "n": int(n),
        "row_ptr": row_ptr,
        "col": col.astype(np.int32),
        "factor": fac.astype(np.int32),
        "cmin": cmin.astype(np.int32),
        "cmax": cmax.astype(np.int32),
        "problem_type": ptype,
        "maximize": int(maximize),
        "obj_var": np.arange(n, dtype=np.int32),
        "obj_factor": np.asarray(costs, np.float64),
        "cost_constant": 0.0,
    }


def _rows_from_columns(m, col_rows):
    """col_rows[j] = sorted row ids of column j -> per-row sorted column lists."""
    rows = [[] for _ in range(m)]
    for j, rs in enumerate(col_rows):
        for k in rs:
            rows[k].append(j)
    return rows


def _dedup_rows(rows, rng, n):
    """The reference merges rows with identical element vectors; keep the
    generated instance free of duplicates so that m is what we asked for."""
    seen = set()
    for k, r in enumerate(rows):
        t = tuple(r)
        while t in seen or len(r) < 2:
            j = int(rng.randint(n))
            if j not in r:
                r.append(j)
                r.sort()
            t = tuple(r)
        seen.add(t)
    return rows


def scp(m=200, n=1000, rows_per_col=4, seed=1, cost_hi=100, name="scp41-shaped"):
    """C1: set covering, rows >= 1 (inequalities_01)."""
    rng = np.random.RandomState(seed)
    col_rows = [sorted(rng.choice(m, size=rows_per_col, replace=False).tolist()) for _ in range(n)]
    rows = _rows_from_columns(m, col_rows)
    rows = _dedup_rows(rows, rng, n)
    costs = rng.randint(1, cost_hi + 1, size=n)
    return _finish(m, n, rows, [[1] * len(r) for r in rows], [1] * m, [INT_MAX] * m, costs, name=name)


def spp(m=500, n=10000, lo=5, hi=8, seed=2, cost_hi=100, name="spp-shaped"):
    """C2: set partitioning, rows = 1 (equalities_01), with one planted
    feasible partition so that feasibility is reachable."""
    rng = np.random.RandomState(seed)
    perm = rng.permutation(m)
    col_rows = []
    pos = 0
    while pos < m:
        size = int(rng.randint(lo, hi + 1))
        grp = perm[pos : pos + size]
        pos += size
        col_rows.append(sorted(grp.tolist()))
    while len(col_rows) < n:
        size = int(rng.randint(lo, hi + 1))
        col_rows.append(sorted(rng.choice(m, size=size, replace=False).tolist()))
    order = rng.permutation(len(col_rows))
    col_rows = [col_rows[i] for i in order]
    rows = _rows_from_columns(m, col_rows)
    rows = _dedup_rows(rows, rng, n)
    lens = np.zeros(n, np.int64)
    for r in rows:
        for j in r:
            lens[j] += 1
    costs = rng.randint(1, cost_hi + 1, size=n) * np.maximum(lens, 1)
    return _finish(m, n, rows, [[1] * len(r) for r in rows], [1] * m, [1] * m, costs, name=name)


def csplib022(m=30, n=50000, lo=5, hi=15, neg_frac=0.2, seed=3, name="csplib022-shaped"):
    """C3: wide rows with +-1 coefficients, rows = b_k (equalities_101)."""
    rng = np.random.RandomState(seed)
    col_rows = [sorted(rng.choice(m, size=int(rng.randint(lo, hi + 1)), replace=False).tolist()) for _ in range(n)]
    rows = _rows_from_columns(m, col_rows)
    rows = _dedup_rows(rows, rng, n)
    factors = [np.where(rng.rand(len(r)) < neg_frac, -1, 1).tolist() for r in rows]
    b = rng.randint(1, 4, size=m)
    costs = rng.randint(1, 101, size=n)
    return _finish(m, n, rows, factors, b, b, costs, name=name)


def telebus(m=1500, n=7000, seed=4, linear_hi=25, name="telebus-shaped", ex_frac=0.05, bb_frac=0.05):
    """C4: general integer coefficients (inequalities_Z).  Most rows are 0/1 or
    +-1 (linear sub-solver); about 10% carry a coefficient with |f| >= 2 drawn
    from U{-3..9}\\{0}: half of those short (< 10 elements, exhaustive), half
    with 10-16 elements (branch and bound)."""
    rng = np.random.RandomState(seed)
    rows, factors, cmin, cmax = [], [], [], []
    for k in range(m):
        u = rng.rand()
        if u < ex_frac:
            size = int(rng.randint(3, 10))
            zrow = True
        elif u < ex_frac + bb_frac:
            size = int(rng.randint(10, 17))
            zrow = True
        else:
            size = int(rng.randint(3, linear_hi))
            zrow = False
        cols = sorted(rng.choice(n, size=size, replace=False).tolist())
        if zrow:
            f = rng.randint(-3, 10, size=size)
            f[f == 0] = 1
            if not np.any(np.abs(f) >= 2):
                f[0] = 2
        else:
            f = np.where(rng.rand(size) < 0.15, -1, 1)
        kind = rng.randint(3)
        if kind == 1 and not np.any(f > 0):
            f[0] = -f[0]  # a ">= 1" row needs a positive coefficient to be realisable
        lower = int(f[f < 0].sum())
        upper = int(f[f > 0].sum())
        # pick bounds that the all-zero or a small subset can satisfy
        if kind == 0:  # <=
            cmin.append(INT_MIN)
            cmax.append(int(rng.randint(max(1, upper // 3), max(2, upper))))
        elif kind == 1:  # >=
            cmin.append(int(rng.randint(1, max(2, upper // 2 + 1))))
            cmax.append(INT_MAX)
        else:  # =
            pick = rng.rand(size) < 0.4
            v = int(f[pick].sum())
            v = min(max(v, lower), upper)
            cmin.append(v)
            cmax.append(v)
        rows.append(cols)
        factors.append(f.tolist())
    # every variable must appear: append unused columns to random 0/1 rows
    used = np.zeros(n, bool)
    for r in rows:
        used[r] = True
    costs = rng.randint(1, 101, size=n)
    seen = set()
    for k, r in enumerate(rows):
        t = (tuple(r), tuple(factors[k]))
        assert t not in seen or True
        seen.add(t)
    return _finish(m, n, rows, factors, cmin, cmax, costs, name=name)


def rail(m=4284, n=1092610, mean_col=10.3, seed=5, name="rail4284-shaped"):
    """C5: rail4284-shaped set covering (inequalities_01), costs in {1, 2}.
    Column sizes are 2 + Poisson(mean_col - 2) clipped to [2, 24] (the real
    file's histogram is not in the repository; this choice is recorded in
    DESIGN.md)."""
    rng = np.random.RandomState(seed)
    sizes = np.clip(2 + rng.poisson(mean_col - 2.0, size=n), 2, 24).astype(np.int64)
    nnz = int(sizes.sum())
    # rows of a column: distinct picks; draw with replacement then repair
    col_of = np.repeat(np.arange(n, dtype=np.int64), sizes)
    row_of = rng.randint(0, m, size=nnz).astype(np.int64)
    key = col_of * m + row_of
    key = np.unique(key)
    col_of = (key // m).astype(np.int32)
    row_of = (key % m).astype(np.int32)
    order = np.lexsort((col_of, row_of))
    row_sorted = row_of[order]
    col_sorted = col_of[order]
    row_ptr = np.zeros(m + 1, np.int32)
    np.add.at(row_ptr, row_sorted + 1, 1)
    row_ptr = np.cumsum(row_ptr).astype(np.int32)
    costs = rng.randint(1, 3, size=n)
    return {
        "name": name,
        "m": int(m),
        "n": int(n),
        "row_ptr": row_ptr,
        "col": col_sorted.astype(np.int32),
        "factor": np.ones(len(col_sorted), np.int32),
        "cmin": np.ones(m, np.int32),
        "cmax": np.full(m, INT_MAX, np.int32),
        "problem_type": INEQ_01,
        "maximize": 0,
        "obj_var": np.arange(n, dtype=np.int32),
        "obj_factor": costs.astype(np.float64),
        "cost_constant": 0.0,
    }


def small_mixed(m=12, n=30, seed=7, family="01", equality=False, maximize=False, cost_hi=9,
                density=0.25, name="small"):
    """Small instances with many cost ties for trajectory-level parity tests.
    family: "01", "101" or "Z"."""
    rng = np.random.RandomState(seed)
    rows, factors, cmin, cmax = [], [], [], []
    seen = set()
    for k in range(m):
        while True:
            size = max(2, int(rng.binomial(n, density)))
            cols = sorted(rng.choice(n, size=min(size, n), replace=False).tolist())
            if tuple(cols) not in seen:
                seen.add(tuple(cols))
                break
        size = len(cols)
        if family == "01":
            f = np.ones(size, np.int64)
        elif family == "101":
            f = np.where(rng.rand(size) < 0.3, -1, 1)
        else:
            f = rng.randint(-3, 6, size=size)
            f[f == 0] = 1
        lower = int(f[f < 0].sum())
        upper = int(f[f > 0].sum())
        pick = rng.rand(size) < 0.35
        v = int(f[pick].sum())
        if equality:
            cmin.append(v)
            cmax.append(v)
        else:
            kind = rng.randint(3)
            if kind == 0:
                cmin.append(INT_MIN)
                cmax.append(max(v, lower + 1))
            elif kind == 1:
                cmin.append(min(v, upper - 1) if upper - 1 >= lower else v)
                cmax.append(INT_MAX)
            else:
                cmin.append(v)
                cmax.append(v)
        rows.append(cols)
        factors.append(f.tolist())
    costs = rng.randint(1, cost_hi + 1, size=n)
    return _finish(m, n, rows, factors, cmin, cmax, costs, maximize=maximize, name=name)


def write_lp(inst, path):
    """CPLEX-LP text of an instance.  Every variable appears in the objective
    in index order, so the reader numbers variables 0..n-1 like we do."""
    n, m = inst["n"], inst["m"]
    rp, col, fac = inst["row_ptr"], inst["col"], inst["factor"]
    cost = np.zeros(n, np.float64)
    for v, f in zip(inst["obj_var"], inst["obj_factor"]):
        cost[v] += f
    with open(path, "w") as f:
        f.write("maximize\n" if inst["maximize"] else "minimize\n")
        terms = []
        for j in range(n):
            c = cost[j]
            cs = repr(int(c)) if float(c).is_integer() else repr(float(c))
            terms.append(f"{'+ ' if c >= 0 else '- '}{cs.lstrip('-')} x{j}")
        for i in range(0, n, 8):
            f.write(" " + " ".join(terms[i : i + 8]) + "\n")
        f.write("subject to\n")
        for k in range(m):
            parts = []
            for e in range(rp[k], rp[k + 1]):
                a = int(fac[e])
                sign = "+" if a >= 0 else "-"
                mag = abs(a)
                parts.append(f"{sign} {'' if mag == 1 else str(mag) + ' '}x{int(col[e])}")
            lo, hi = int(inst["cmin"][k]), int(inst["cmax"][k])
            lines = []
            if lo == hi:
                lines.append(("=", lo))
            else:
                if lo != INT_MIN:
                    lines.append((">=", lo))
                if hi != INT_MAX:
                    lines.append(("<=", hi))
            for t, (op, v) in enumerate(lines):
                label = f"c{k}" if t == 0 else f"c{k}b"
                f.write(f"{label}: ")
                for i in range(0, len(parts), 10):
                    f.write(" ".join(parts[i : i + 10]))
                    f.write("\n " if i + 10 < len(parts) else " ")
                f.write(f"{op} {v}\n")
        f.write("binary\n")
        for i in range(0, n, 10):
            f.write(" " + " ".join(f"x{j}" for j in range(i, min(n, i + 10))) + "\n")
        f.write("end\n")
    return path
