"""ctypes bindings for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

* ``liboracle.so``      -- plain-C restatement of the hot path (bx_oracle.c)
* ``_ref/libbxref.so``  -- the unmodified reference compiled by oracle/Makefile
                           (present in the build container; travels to the GPU
                           box as a prebuilt file)

Only tests/, ``__graft_entry__.smoke()`` and the cpu_baseline / ``--impl
reference`` legs of bench.py import this module.  The product package
``baryonyx_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = "/root/reference"


class Params(C.Structure):
    """Mirror of baryonyx::solver_parameters (lib/include/baryonyx/core:294-462).

    Same layout for ``ox_params`` (bx_oracle.h), ``bxref_params`` (bxref.h) and
    ``bxg_params`` (include/bxg.h).
    """

    _fields_ = [
        ("time_limit", C.c_double),
        ("theta", C.c_double),
        ("delta", C.c_double),
        ("kappa_min", C.c_double),
        ("kappa_step", C.c_double),
        ("kappa_max", C.c_double),
        ("alpha", C.c_double),
        ("pushing_k_factor", C.c_double),
        ("pushing_objective_amplifier", C.c_double),
        ("init_policy_random", C.c_double),
        ("init_crossover_bastert_insertion", C.c_double),
        ("init_crossover_solution_selection_mean", C.c_double),
        ("init_crossover_solution_selection_stddev", C.c_double),
        ("init_mutation_variable_mean", C.c_double),
        ("init_mutation_variable_stddev", C.c_double),
        ("init_mutation_value_mean", C.c_double),
        ("init_mutation_value_stddev", C.c_double),
        ("init_kappa_improve_start", C.c_double),
        ("init_kappa_improve_increase", C.c_double),
        ("init_kappa_improve_stop", C.c_double),
        ("w", C.c_double),
        ("seed", C.c_uint32),
        ("thread", C.c_int32),
        ("limit", C.c_int64),
        ("print_level", C.c_int32),
        ("pushes_limit", C.c_int32),
        ("pushing_iteration_limit", C.c_int32),
        ("init_population_size", C.c_int32),
        ("pre_order", C.c_int32),
        ("order", C.c_int32),
        ("float_type", C.c_int32),
        ("init_policy", C.c_int32),
        ("cost_norm", C.c_int32),
        ("mode", C.c_int32),
        ("preprocessor", C.c_int32),
        ("observer", C.c_int32),
        ("storage", C.c_int32),
        ("solver", C.c_int32),
        ("debug", C.c_int32),
    ]

    def copy(self) -> "Params":
        q = Params()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(Params))
        return q

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


ORDERS = {
    "none": 0,
    "reversing": 1,
    "random-sorting": 2,
    "infeasibility-decr": 3,
    "infeasibility-incr": 4,
    "lagrangian-decr": 5,
    "lagrangian-incr": 6,
    "pi-sign-change": 7,
    "cycle": 8,
}

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")


def _opt(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def build_oracle(force: bool = False) -> str:
    """Compile liboracle.so (gcc) if missing or stale; returns its path."""
    so = os.path.join(HERE, "liboracle.so")
    src = [os.path.join(HERE, f) for f in ("bx_oracle.c", "bx_oracle.h")]
    if force or not os.path.exists(so) or any(
        os.path.getmtime(s) > os.path.getmtime(so) for s in src
    ):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"])
    return so


def build_ref(adapter: bool = True) -> str | None:
    """Compile _ref/libbxref*.so from /root/reference when it is present.
    adapter=False: only the three builds of the reference itself (they do not
    need the engine's library, so they can compile while nvcc runs); the
    default also links _ref/libbxref_b200.so (integration/itm_b200.cpp), which
    needs baryonyx_b200/libbaryonyx_b200.so."""
    so = os.path.join(HERE, "_ref", "libbxref.so")
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "lib", "src")):
        targets = ["ref"] if adapter else ["_ref/libbxref.so", "_ref/libbxref_fast.so", "_ref/libbxref_o3.so"]
        subprocess.check_call(["make", "-C", HERE, "-j", str(os.cpu_count() or 4)] + targets)
    return so if os.path.exists(so) else None


# --------------------------------------------------------------------------
# liboracle.so
# --------------------------------------------------------------------------
_oracle = None


def oracle():
    global _oracle
    if _oracle is not None:
        return _oracle
    lib = C.CDLL(build_oracle())
    lib.ox_default_params.argtypes = [C.POINTER(Params)]
    lib.ox_clamp_params.argtypes = [C.POINTER(Params), C.POINTER(Params)]
    lib.ox_minstd_next.argtypes = [C.POINTER(C.c_uint32)]
    lib.ox_minstd_next.restype = C.c_uint32
    lib.ox_bernoulli.argtypes = [C.c_double, C.POINTER(C.c_uint32)]
    lib.ox_uniform_u64.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32)]
    lib.ox_uniform_u64.restype = C.c_uint64
    lib.ox_shuffle_int.argtypes = [_i32p, C.c_int, C.POINTER(C.c_uint32)]
    lib.ox_sort_rc.argtypes = [_f64p, _i32p, C.c_int, C.c_int]
    lib.ox_sort_pairs.argtypes = [_i32p, _i32p, C.c_int, C.c_int]
    lib.ox_normal.argtypes = [C.c_double, C.c_double, C.c_int, _f64p, C.POINTER(C.c_uint32)]
    lib.ox_problem_create.argtypes = [
        C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p, _i32p, C.c_int, C.c_int,
        C.c_int, _i32p, _f64p, C.c_double,
    ]
    lib.ox_problem_create.restype = C.c_void_p
    lib.ox_problem_destroy.argtypes = [C.c_void_p]
    lib.ox_problem_set_tail_value.argtypes = [C.c_void_p, C.c_double]
    lib.ox_which_problem_type.argtypes = [C.c_int, _i32p, C.c_int]
    lib.ox_run_begin.argtypes = [C.c_void_p, C.POINTER(Params), C.c_uint32]
    lib.ox_run_begin.restype = C.c_void_p
    lib.ox_run_end.argtypes = [C.c_void_p]
    lib.ox_run_info.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.ox_run_sweep.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
    lib.ox_run_state.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    lib.ox_run_set_x.argtypes = [C.c_void_p, _u64p]
    lib.ox_run_set_rng.argtypes = [C.c_void_p, C.c_uint32]
    lib.ox_run_solve.argtypes = [C.c_void_p, C.c_double, C.c_int, _u64p, C.POINTER(C.c_double),
                                 C.POINTER(C.c_int), C.POINTER(C.c_int64)]
    lib.ox_run_objective.argtypes = [C.c_void_p]
    lib.ox_run_objective.restype = C.c_double
    lib.ox_run_updates.argtypes = [C.c_void_p]
    lib.ox_run_updates.restype = C.c_int64
    lib.ox_solve.argtypes = [
        C.c_void_p, C.POINTER(Params), C.c_uint32, _i8p, C.POINTER(C.c_double),
        C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
    ]
    lib.ox_opt_trace_run.argtypes = [C.c_void_p, C.POINTER(Params), C.c_uint32, C.c_int]
    lib.ox_opt_trace_run.restype = C.c_void_p
    lib.ox_opt_trace_free.argtypes = [C.c_void_p]
    lib.ox_opt_trace_rounds.argtypes = [C.c_void_p]
    lib.ox_opt_trace_round.argtypes = [C.c_void_p, C.c_int, _u64p, C.POINTER(C.c_uint32), C.POINTER(C.c_double),
                                       _f64p, _f64p]
    lib.ox_opt_trace_member.argtypes = [C.c_void_p, C.c_int, C.c_int, _u64p, C.POINTER(C.c_double),
                                        C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]
    lib.ox_opt_trace_sweeps.argtypes = [C.c_void_p]
    lib.ox_opt_trace_sweep_log.argtypes = [C.c_void_p, _i32p, _i32p, _f64p, _i32p]
    lib.ox_bb_solve.argtypes = [C.c_int, _f64p, _i32p, _i32p, C.c_int, C.c_int, C.c_int]
    lib.ox_exhaustive_solve.argtypes = [C.c_int, _i32p, _f64p, _i32p, _i32p, C.c_int, C.c_int, C.c_int]
    _oracle = lib
    return lib


def default_params() -> Params:
    p = Params()
    oracle().ox_default_params(C.byref(p))
    return p


def clamp_params(p: Params) -> Params:
    q = Params()
    oracle().ox_clamp_params(C.byref(p), C.byref(q))
    return q


_GRID_NAMES = ("theta", "delta", "kappa_min", "kappa_step", "init_policy_random")


def manual_grid(p: Params) -> dict:
    """ox_manual_grid: the manual auto-tuner's 3125 parameter sets + its tables."""
    arrs = [np.zeros(3125, np.float64) for _ in _GRID_NAMES]
    tables = np.zeros((5, 5), np.float64)
    f = oracle().ox_manual_grid
    f.argtypes = [C.POINTER(Params)] + [_f64p] * 5 + [_f64p]
    n = f(C.byref(p), *arrs, tables.reshape(-1))
    assert n == 3125
    out = dict(zip(_GRID_NAMES, arrs))
    out["tables"] = tables
    return out


def ref_manual_grid(p: Params) -> dict:
    """bxref_manual_grid: the same from the compiled reference's manual_course."""
    arrs = [np.zeros(4000, np.float64) for _ in _GRID_NAMES]
    tables = np.zeros((5, 5), np.float64)
    f = ref(False).bxref_manual_grid
    f.argtypes = [C.POINTER(Params)] + [_f64p] * 5 + [C.c_int, _f64p]
    n = f(C.byref(p), *arrs, 4000, tables.reshape(-1))
    out = {k: a[:n] for k, a in zip(_GRID_NAMES, arrs)}
    out["tables"] = tables
    return out


class _RunBase:
    """Common state accessors for step-level runs (oracle and reference)."""

    def __init__(self, m, n, nnz):
        self.m, self.n, self.nnz = m, n, nnz
        self.words = n // 64 + 1

    def state(self, order_cap=None):
        x = np.zeros(self.words, np.uint64)
        pi = np.zeros(self.m, np.float64)
        P = np.zeros(self.nnz, np.float64)
        rng = C.c_uint32(0)
        cap = order_cap or (2 * max(self.m, self.n) + 2)
        ol = np.zeros(cap, np.int32)
        on = C.c_int(0)
        self._state(_opt(x), _opt(pi), _opt(P), C.cast(C.byref(rng), C.c_void_p),
                    _opt(ol), C.cast(C.byref(on), C.c_void_p), None)
        return {"x": x, "pi": pi, "P": P, "rng": rng.value, "order": ol, "order_len": on.value}

    def norm_cost(self):
        c = np.zeros(self.n, np.float64)
        self._state(None, None, None, None, None, None, _opt(c))
        return c


class OracleProblem:
    def __init__(self, inst):
        """inst: dict with m, n, row_ptr, col, factor, cmin, cmax, problem_type,
        maximize, obj_var, obj_factor, cost_constant (see baryonyx_b200.synth)."""
        self.inst = inst
        self.lib = oracle()
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        self.h = self.lib.ox_problem_create(
            inst["m"], inst["n"], i32(inst["row_ptr"]), i32(inst["col"]), i32(inst["factor"]),
            i32(inst["cmin"]), i32(inst["cmax"]), int(inst["problem_type"]), int(inst["maximize"]),
            len(inst["obj_var"]), i32(inst["obj_var"]),
            np.ascontiguousarray(inst["obj_factor"], np.float64), float(inst["cost_constant"]),
        )
        if not self.h:
            raise RuntimeError("ox_problem_create failed (unrealisable constraint?)")
        self.m, self.n, self.nnz = inst["m"], inst["n"], int(inst["row_ptr"][-1])

    def set_tail_value(self, v: float):
        self.lib.ox_problem_set_tail_value(self.h, float(v))

    def close(self):
        if self.h:
            self.lib.ox_problem_destroy(self.h)
            self.h = None

    def run(self, params: Params, seed: int) -> "OracleRun":
        return OracleRun(self, params, seed)

    def optimize_trace(self, params: Params, seed: int, rounds: int):
        """optimize_functor at thread = 1 for `rounds` restarts (ox_opt_trace_run);
        `params` clamped.  Same arrays as RefProblem.optimize_trace."""
        t = self.lib.ox_opt_trace_run(self.h, C.byref(params), seed, rounds)
        if not t:
            raise RuntimeError("ox_opt_trace_run failed")
        try:
            return read_opt_trace(self.lib, "ox_opt_trace", t, self.m, self.n, self.nnz,
                                  params.init_population_size)
        finally:
            self.lib.ox_opt_trace_free(t)

    def solve(self, params: Params, seed: int):
        x = np.zeros(self.n, np.int8)
        value = C.c_double()
        rem = C.c_int()
        loop = C.c_int64()
        upd = C.c_int64()
        swp = C.c_int64()
        st = self.lib.ox_solve(self.h, C.byref(params), seed, x, C.byref(value), C.byref(rem),
                               C.byref(loop), C.byref(upd), C.byref(swp))
        return {"status": st, "x": x, "value": value.value, "remaining": rem.value,
                "loop": loop.value, "updates": upd.value, "sweeps": swp.value}


class OracleRun(_RunBase):
    def __init__(self, pb: OracleProblem, params: Params, seed: int):
        super().__init__(pb.m, pb.n, pb.nnz)
        self.lib = pb.lib
        self.h = self.lib.ox_run_begin(pb.h, C.byref(params), seed)
        if not self.h:
            raise RuntimeError("ox_run_begin failed (unsupported cost_norm?)")
        self._state = lambda *a: self.lib.ox_run_state(self.h, *a)

    def info(self):
        d, t = C.c_double(), C.c_double()
        self.lib.ox_run_info(self.h, C.byref(d), C.byref(t))
        return d.value, t.value

    def sweep(self, kappa, push=0, obj_amp=0.0):
        return self.lib.ox_run_sweep(self.h, push, kappa, obj_amp)

    def set_x(self, words):
        self.lib.ox_run_set_x(self.h, np.ascontiguousarray(words, np.uint64))

    def set_rng(self, s):
        self.lib.ox_run_set_rng(self.h, s)

    def objective(self):
        return self.lib.ox_run_objective(self.h)

    def solve(self, kappa_start: float, optimizer: bool = False):
        """The solver_functor loop (or one optimizer call) on this run as it stands."""
        bx = np.zeros(self.n // 64 + 1, np.uint64)
        v, r, l = C.c_double(), C.c_int(), C.c_int64()
        st = self.lib.ox_run_solve(self.h, kappa_start, int(optimizer), bx, C.byref(v), C.byref(r),
                                   C.byref(l))
        return {"status": st, "best_x": bx, "value": v.value, "remaining": r.value, "loop": l.value}

    def updates(self):
        return self.lib.ox_run_updates(self.h)

    def close(self):
        if self.h:
            self.lib.ox_run_end(self.h)
            self.h = None


# --------------------------------------------------------------------------
# _ref/libbxref.so
# --------------------------------------------------------------------------
_ref = {}


def _ref_name(fast) -> str:
    """fast: False = parity build (contract checks on), True = the same without
    the checks, "o3" = the reference's release flags (speed baseline only)."""
    return {False: "libbxref.so", True: "libbxref_fast.so", "o3": "libbxref_o3.so"}[fast]


def ref_available(fast=False) -> bool:
    if fast == "o3" and not _cpu_has("avx2", "fma", "bmi2"):
        return False  # built for x86-64-v3
    return os.path.exists(os.path.join(HERE, "_ref", _ref_name(fast)))


def _cpu_has(*flags) -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    have = set(line.split(":", 1)[1].split())
                    return all(fl in have for fl in flags)
    except OSError:
        pass
    return False


def ref_b200():
    """oracle/_ref/libbxref_b200.so: integration/itm_b200.cpp compiled against the
    reference's headers (the only harness library that links the engine)."""
    if "b200" not in _ref:
        ref(False)
        lib = C.CDLL(os.path.join(HERE, "_ref", "libbxref_b200.so"))
        lib.bxref_solve_b200.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, _i8p, C.POINTER(C.c_double),
                                         C.POINTER(C.c_int), C.POINTER(C.c_int64)]
        lib.bxref_optimize_b200.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_double, C.c_int64, C.c_int,
                                            _i8p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                            C.POINTER(C.c_int * 3)]
        _ref["b200"] = lib
    return _ref["b200"]


def ref(fast=False):
    if fast in _ref:
        return _ref[fast]
    path = os.path.join(HERE, "_ref", _ref_name(fast))
    lib = C.CDLL(path)
    lib.bxref_default_params.argtypes = [C.POINTER(Params)]
    lib.bxref_open.argtypes = [C.c_char_p, C.POINTER(Params)]
    lib.bxref_open.restype = C.c_void_p
    lib.bxref_close.argtypes = [C.c_void_p]
    lib.bxref_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 5 + [C.POINTER(C.c_double), C.POINTER(C.c_int)]
    lib.bxref_get_params.argtypes = [C.c_void_p, C.POINTER(Params)]
    lib.bxref_export_rows.argtypes = [C.c_void_p, _i32p, _i32p, _i32p, _i32p, _i32p]
    lib.bxref_export_objective.argtypes = [C.c_void_p, _i32p, _f64p]
    lib.bxref_run_begin.argtypes = [C.c_void_p, C.c_uint32]
    lib.bxref_run_begin.restype = C.c_void_p
    lib.bxref_run_end.argtypes = [C.c_void_p]
    lib.bxref_run_info.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.bxref_run_sweep.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
    lib.bxref_run_tail_value.argtypes = [C.c_void_p]
    lib.bxref_run_tail_value.restype = C.c_double
    lib.bxref_run_state.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    lib.bxref_run_set_x.argtypes = [C.c_void_p, _u64p]
    lib.bxref_run_set_rng.argtypes = [C.c_void_p, C.c_uint32]
    lib.bxref_run_objective.argtypes = [C.c_void_p]
    lib.bxref_run_objective.restype = C.c_double
    lib.bxref_solve.argtypes = [C.c_void_p, C.c_uint32, _i8p, C.POINTER(C.c_double), C.POINTER(C.c_int),
                                C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.bxref_solve_public.argtypes = [C.c_void_p, C.c_uint32, _i8p, C.POINTER(C.c_double)]
    lib.bxref_optimize.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_double, _i8p, C.POINTER(C.c_double),
                                   C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                   C.POINTER(C.c_double)]
    lib.bxref_opt_trace.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
    lib.bxref_opt_trace.restype = C.c_void_p
    lib.bxref_opt_trace_free.argtypes = [C.c_void_p]
    lib.bxref_opt_trace_rounds.argtypes = [C.c_void_p]
    lib.bxref_opt_trace_round.argtypes = [C.c_void_p, C.c_int, _u64p, C.POINTER(C.c_uint32), C.POINTER(C.c_double),
                                          _f64p, _f64p]
    lib.bxref_opt_trace_member.argtypes = [C.c_void_p, C.c_int, C.c_int, _u64p, C.POINTER(C.c_double),
                                           C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]
    lib.bxref_opt_trace_sweeps.argtypes = [C.c_void_p]
    lib.bxref_opt_trace_sweep_log.argtypes = [C.c_void_p, _i32p, _i32p, _f64p, _i32p]
    lib.bxref_std_sort.argtypes = [_f64p, _i32p, C.c_int, C.c_int]
    lib.bxref_std_sort_pairs.argtypes = [_i32p, _i32p, C.c_int, C.c_int]
    lib.bxref_std_shuffle.argtypes = [_i32p, C.c_int, C.POINTER(C.c_uint32)]
    lib.bxref_std_bernoulli.argtypes = [C.c_double, C.POINTER(C.c_uint32)]
    lib.bxref_std_uniform_u64.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32)]
    lib.bxref_std_uniform_u64.restype = C.c_uint64
    lib.bxref_std_normal.argtypes = [C.c_double, C.c_double, C.c_int, _f64p, C.POINTER(C.c_uint32)]
    lib.bxref_generate_seed.argtypes = [C.c_uint64, C.c_uint, np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")]
    lib.bxref_open_arrays.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p, _i32p, C.c_int, C.c_int, _i32p,
                                      _f64p, C.c_double, C.POINTER(Params)]
    lib.bxref_open_arrays.restype = C.c_void_p
    lib.bxref_std_pow.argtypes = [C.c_double, C.c_double]
    lib.bxref_std_pow.restype = C.c_double
    lib.bxref_bb_solve.argtypes = [C.c_int, _f64p, _i32p, _i32p, C.c_int, C.c_int, C.c_int]
    lib.bxref_exhaustive_solve.argtypes = [C.c_int, _i32p, _f64p, _i32p, _i32p, C.c_int, C.c_int, C.c_int]
    _ref[fast] = lib
    return lib


def read_opt_trace(lib, prefix, t, m, n, nnz, size):
    """Arrays of an optimizer trace (reference harness or oracle, same layout)."""
    f = lambda name: getattr(lib, prefix + "_" + name)
    words = n // 64 + 1
    R = f("rounds")(t)
    out = {"x": np.zeros((R, words), np.uint64), "rng": np.zeros(R, np.uint32), "kappa": np.zeros(R),
           "pi": np.zeros((R, m)), "P": np.zeros((R, nnz)),
           "pop_x": np.zeros((R, size, words), np.uint64), "pop_value": np.zeros((R, size)),
           "pop_remaining": np.zeros((R, size), np.int32), "pop_hash": np.zeros((R, size), np.uint64),
           "pop_loop": np.zeros((R, size), np.int64)}
    for r in range(R):
        rng, kappa = C.c_uint32(), C.c_double()
        f("round")(t, r, out["x"][r], C.byref(rng), C.byref(kappa), out["pi"][r], out["P"][r])
        out["rng"][r], out["kappa"][r] = rng.value, kappa.value
        for i in range(size):
            v, rem, h, lp = C.c_double(), C.c_int(), C.c_uint64(), C.c_int64()
            f("member")(t, r, i, out["pop_x"][r, i], C.byref(v), C.byref(rem), C.byref(h), C.byref(lp))
            out["pop_value"][r, i], out["pop_remaining"][r, i] = v.value, rem.value
            out["pop_hash"][r, i], out["pop_loop"][r, i] = h.value, lp.value
    S = f("sweeps")(t)
    out["sweep_round"] = np.zeros(S, np.int32)
    out["sweep_push"] = np.zeros(S, np.int32)
    out["sweep_kappa"] = np.zeros(S)
    out["sweep_len"] = np.zeros(S, np.int32)
    if S:
        f("sweep_log")(t, out["sweep_round"], out["sweep_push"], out["sweep_kappa"], out["sweep_len"])
    return out


class RefProblem:
    """LP file -> reference parser -> merged constraints."""

    def __init__(self, lp_path: str | None, params: Params, fast=False, inst: dict | None = None,
                 preprocess: bool = False, affect=None):
        """From an LP file (the reference's parser), or from an instance dict
        (bxref_open_arrays: raw_problem filled directly, for instances whose LP
        text the reference's parser takes minutes to read)."""
        self.lib = ref(fast)
        if inst is not None:
            i32 = lambda a: np.ascontiguousarray(a, np.int32)
            self.h = self.lib.bxref_open_arrays(
                inst["m"], inst["n"], i32(inst["row_ptr"]), i32(inst["col"]), i32(inst["factor"]),
                i32(inst["cmin"]), i32(inst["cmax"]), int(inst["maximize"]), len(inst["obj_var"]),
                i32(inst["obj_var"]), np.ascontiguousarray(inst["obj_factor"], np.float64),
                float(inst["cost_constant"]), C.byref(params))
        elif preprocess:
            # baryonyx::preprocess, then baryonyx::affect for every (variable, value)
            # (lib/src/preprocessor.cpp:876-903).  The reference leaves a
            # "preprocessed.lp" in the working directory: run from a scratch one.
            f = self.lib.bxref_open_pp
            f.argtypes = [C.c_char_p, C.POINTER(Params), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                          C.c_char_p, C.c_int]
            f.restype = C.c_void_p
            aff = list(affect or [])
            va = (C.c_int * max(1, len(aff)))(*[int(a[0]) for a in aff])
            vb = (C.c_int * max(1, len(aff)))(*[int(a[1]) for a in aff])
            err = C.create_string_buffer(512)
            import tempfile
            cwd = os.getcwd()
            with tempfile.TemporaryDirectory() as scratch:
                os.chdir(scratch)
                try:
                    self.h = f(os.path.abspath(os.path.join(cwd, lp_path)).encode(), C.byref(params), len(aff),
                               va, vb, err, 512)
                finally:
                    os.chdir(cwd)
            if not self.h:
                raise RuntimeError(f"reference preprocessor: {err.value.decode()}")
        else:
            self.h = self.lib.bxref_open(lp_path.encode(), C.byref(params))
        if not self.h:
            raise RuntimeError(f"reference could not read {lp_path}")
        v = [C.c_int() for _ in range(5)]
        cc, nobj = C.c_double(), C.c_int()
        self.lib.bxref_info(self.h, *[C.byref(a) for a in v], C.byref(cc), C.byref(nobj))
        self.m, self.n, self.nnz, self.problem_type, self.maximize = [a.value for a in v]
        self.cost_constant, self.n_obj = cc.value, nobj.value

    def params(self) -> Params:
        q = Params()
        self.lib.bxref_get_params(self.h, C.byref(q))
        return q

    def instance(self) -> dict:
        """Merged constraints + objective as arrays (the oracle / engine input)."""
        row_ptr = np.zeros(self.m + 1, np.int32)
        col = np.zeros(max(self.nnz, 1), np.int32)
        fac = np.zeros(max(self.nnz, 1), np.int32)
        cmin = np.zeros(self.m, np.int32)
        cmax = np.zeros(self.m, np.int32)
        self.lib.bxref_export_rows(self.h, row_ptr, col, fac, cmin, cmax)
        ov = np.zeros(max(self.n_obj, 1), np.int32)
        of = np.zeros(max(self.n_obj, 1), np.float64)
        self.lib.bxref_export_objective(self.h, ov, of)
        return {
            "m": self.m, "n": self.n, "row_ptr": row_ptr, "col": col[: self.nnz], "factor": fac[: self.nnz],
            "cmin": cmin, "cmax": cmax, "problem_type": self.problem_type, "maximize": self.maximize,
            "obj_var": ov[: self.n_obj], "obj_factor": of[: self.n_obj], "cost_constant": self.cost_constant,
        }

    def variables(self):
        """(names, [(min, max, variable_type)]) of the parsed problem."""
        f = self.lib.bxref_variable
        f.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 3
        f.restype = C.c_char_p
        names, bounds = [], []
        for i in range(self.n):
            lo, hi, ty = C.c_int(), C.c_int(), C.c_int()
            names.append(f(self.h, i, C.byref(lo), C.byref(hi), C.byref(ty)).decode())
            bounds.append((lo.value, hi.value, ty.value))
        return names, bounds

    def affected(self):
        """[(name, value)] of problem::affected_vars (the preprocessor's fixed variables)."""
        f = self.lib.bxref_affected
        f.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        f.restype = C.c_char_p
        self.lib.bxref_affected_count.argtypes = [C.c_void_p]
        out = []
        for i in range(self.lib.bxref_affected_count(self.h)):
            v = C.c_int()
            out.append((f(self.h, i, C.byref(v)).decode(), v.value))
        return out

    def run(self, seed: int) -> "RefRun":
        return RefRun(self, seed)

    def solve_b200(self, seed: int, restarts: int = 1, device: int = 0):
        """baryonyx::itm::solve_b200 (integration/itm_b200.cpp): the reference's
        problem and result types over the engine's C ABI."""
        x = np.zeros(self.n, np.int8)
        value, rem, loop = C.c_double(), C.c_int(), C.c_int64()
        st = ref_b200().bxref_solve_b200(self.h, seed, restarts, device, x, C.byref(value), C.byref(rem),
                                         C.byref(loop))
        return {"status": st, "x": x, "value": value.value, "remaining": rem.value, "loop": loop.value}

    def optimize_b200(self, seed: int, workers: int = 1, device: int = 0, time_limit: float = -1.0,
                      max_rounds: int = 0, storage: int = 0):
        """baryonyx::itm::optimize_b200 (integration/itm_b200.cpp): the reference's
        optimize over the engine, its context callbacks counted."""
        x = np.zeros(self.n, np.int8)
        value, rem, nsol = C.c_double(), C.c_int(), C.c_int()
        calls = (C.c_int * 3)()
        st = ref_b200().bxref_optimize_b200(self.h, seed, workers, device, time_limit, max_rounds, storage, x,
                                            C.byref(value), C.byref(rem), C.byref(nsol), C.byref(calls))
        return {"status": st, "x": x, "value": value.value, "remaining": rem.value, "solutions": nsol.value,
                "calls": {"start": calls[0], "update": calls[1], "finish": calls[2]}}

    def solve(self, seed: int):
        x = np.zeros(self.n, np.int8)
        value, rem = C.c_double(), C.c_int()
        upd, swp = C.c_int64(), C.c_int64()
        st = self.lib.bxref_solve(self.h, seed, x, C.byref(value), C.byref(rem), C.byref(upd), C.byref(swp))
        return {"status": st, "x": x, "value": value.value, "remaining": rem.value,
                "updates": upd.value, "sweeps": swp.value}

    def solve_public(self, seed: int):
        x = np.zeros(self.n, np.int8)
        value = C.c_double()
        st = self.lib.bxref_solve_public(self.h, seed, x, C.byref(value))
        return {"status": st, "x": x, "value": value.value}

    def optimize(self, seed: int, threads: int, time_limit: float):
        x = np.zeros(self.n, np.int8)
        value, rem = C.c_double(), C.c_int()
        upd, swp, sec = C.c_int64(), C.c_int64(), C.c_double()
        st = self.lib.bxref_optimize(self.h, seed, threads, time_limit, x, C.byref(value), C.byref(rem),
                                     C.byref(upd), C.byref(swp), C.byref(sec))
        ff, rs = C.c_double(), C.c_int64()
        self.lib.bxref_optimize_report.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        self.lib.bxref_optimize_report(C.byref(ff), C.byref(rs))
        return {"status": st, "x": x, "value": value.value, "remaining": rem.value,
                "updates": upd.value, "sweeps": swp.value, "seconds": sec.value,
                "first_feasible_s": None if ff.value < 0 else ff.value, "restarts": rs.value}

    def optimize_trace(self, seed: int, rounds: int, population_size: int):
        """The reference's optimize_functor at thread = 1, stopped after `rounds`
        restarts (bxref_opt_trace): per-restart snapshots (x, engine, kappa, pi, P,
        population in sorted order) and the sweep log."""
        t = self.lib.bxref_opt_trace(self.h, seed, rounds)
        try:
            return read_opt_trace(self.lib, "bxref_opt_trace", t, self.m, self.n, self.nnz, population_size)
        finally:
            self.lib.bxref_opt_trace_free(t)

    def close(self):
        if self.h:
            self.lib.bxref_close(self.h)
            self.h = None


class RefRun(_RunBase):
    def __init__(self, pb: RefProblem, seed: int):
        super().__init__(pb.m, pb.n, pb.nnz)
        self.lib = pb.lib
        self.h = self.lib.bxref_run_begin(pb.h, seed)
        self._state = lambda *a: self.lib.bxref_run_state(self.h, *a)

    def info(self):
        d, t = C.c_double(), C.c_double()
        self.lib.bxref_run_info(self.h, C.byref(d), C.byref(t))
        return d.value, t.value

    def sweep(self, kappa, push=0, obj_amp=0.0):
        return self.lib.bxref_run_sweep(self.h, push, kappa, obj_amp)

    def tail_value(self):
        return self.lib.bxref_run_tail_value(self.h)

    def set_x(self, words):
        self.lib.bxref_run_set_x(self.h, np.ascontiguousarray(words, np.uint64))

    def set_rng(self, s):
        self.lib.bxref_run_set_rng(self.h, s)

    def objective(self):
        return self.lib.bxref_run_objective(self.h)

    def close(self):
        if self.h:
            self.lib.bxref_run_end(self.h)
            self.h = None


# --------------------------------------------------------------------------
# optimizer population: the same Python face over the
# compiled reference (prefix bxref_pop_) and the C restatement (prefix ox_pop_)
# --------------------------------------------------------------------------
_u64arr = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")


def _bind_population(lib, pre, hash_name):
    if getattr(lib, "_pop_bound", False):
        return
    f = lambda name: getattr(lib, pre + name)
    f("create").restype = C.c_void_p
    f("destroy").argtypes = [C.c_void_p]
    f("size").argtypes = [C.c_void_p]
    f("member").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int),
                            C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]
    f("ctx_create").argtypes = [C.c_void_p, C.c_uint, C.c_uint64]
    f("ctx_create").restype = C.c_void_p
    f("ctx_destroy").argtypes = [C.c_void_p]
    f("ctx_rng").argtypes = [C.c_void_p]
    f("ctx_rng").restype = C.c_uint32
    f("reinit").argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, _u64arr]
    f("reinit").restype = C.c_double
    f("mutation").argtypes = [C.c_void_p, C.c_void_p, _u64arr]
    f("crossover").argtypes = [C.c_void_p, C.c_void_p, _u64arr]
    f("try_update").argtypes = [C.c_void_p, C.c_void_p, _u64arr, C.c_double, C.c_int64]
    f("try_advance").argtypes = [C.c_void_p, C.c_void_p, _u64arr, C.c_int, C.c_int64]
    h = getattr(lib, hash_name)
    h.argtypes = [C.c_int, _u64arr]
    h.restype = C.c_uint64
    lib._pop_bound = True


class Population:
    """best_solution_recorder + local_contexts (itm-optimizer-common.hpp:52-594)."""

    def __init__(self, problem, master_seed: int, threads: int = 1, params: Params | None = None):
        self.n = problem.n
        self.words = self.n // 64 + 1
        rng = C.c_uint32()
        if isinstance(problem, RefProblem):
            self.lib, self.pre, hn = problem.lib, "bxref_pop_", "bxref_bit_array_hash"
            _bind_population(self.lib, self.pre, hn)
            self.lib.bxref_pop_create.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_uint32)]
            self.h = self.lib.bxref_pop_create(problem.h, master_seed, threads, C.byref(rng))
        else:
            self.lib, self.pre, hn = problem.lib, "ox_pop_", "ox_bit_array_hash"
            _bind_population(self.lib, self.pre, hn)
            self.lib.ox_pop_create.argtypes = [C.c_void_p, C.POINTER(Params), C.c_uint64, C.c_int,
                                               C.POINTER(C.c_uint32)]
            self.h = self.lib.ox_pop_create(problem.h, C.byref(params), master_seed, threads, C.byref(rng))
        self._hash = getattr(self.lib, hn)
        self.rng_after_ctor = rng.value
        self.ctxs = {}

    def _f(self, name):
        return getattr(self.lib, self.pre + name)

    def size(self) -> int:
        return self._f("size")(self.h)

    def member(self, i: int):
        x = np.zeros(self.words, np.uint64)
        v, r, h, l = C.c_double(), C.c_int(), C.c_uint64(), C.c_int64()
        self._f("member")(self.h, i, x.ctypes.data_as(C.c_void_p), C.byref(v), C.byref(r), C.byref(h),
                          C.byref(l))
        return {"x": x, "value": v.value, "remaining": r.value, "hash": h.value, "loop": l.value}

    def members(self):
        return [self.member(i) for i in range(self.size())]

    def ctx(self, thread_id: int, seed: int):
        c = self._f("ctx_create")(self.h, thread_id, seed)
        self.ctxs[thread_id] = c
        return c

    def ctx_rng(self, thread_id: int) -> int:
        return self._f("ctx_rng")(self.ctxs[thread_id])

    def set_ctx_rng(self, thread_id: int, state: int):
        """(oracle only) the worker's engine continues from `state`."""
        f = self.lib.ox_pop_ctx_set_rng
        f.argtypes = [C.c_void_p, C.c_uint32]
        f(self.ctxs[thread_id], state)

    def reinit(self, thread_id: int, kappa_min: float, kappa_max: float, x: np.ndarray) -> float:
        return self._f("reinit")(self.h, self.ctxs[thread_id], kappa_min, kappa_max, x)

    def mutation(self, thread_id: int, x: np.ndarray):
        self._f("mutation")(self.h, self.ctxs[thread_id], x)

    def crossover(self, thread_id: int, x: np.ndarray):
        self._f("crossover")(self.h, self.ctxs[thread_id], x)

    def try_update(self, thread_id: int, x: np.ndarray, value: float, loop: int):
        self._f("try_update")(self.h, self.ctxs[thread_id], x, value, loop)

    def try_advance(self, thread_id: int, x: np.ndarray, remaining: int, loop: int):
        self._f("try_advance")(self.h, self.ctxs[thread_id], x, remaining, loop)

    def hash(self, x: np.ndarray) -> int:
        return self._hash(self.n, x)

    def close(self):
        for c in self.ctxs.values():
            self._f("ctx_destroy")(c)
        self.ctxs = {}
        if self.h:
            self._f("destroy")(self.h)
            self.h = None


def ref_preprocess(lp_path: str, params: "Params", affect=None):
    """The compiled reference's preprocess (+ affect chain) on an LP file, run in
    a forked child: a contract failure inside the reference (bx_reach, bx_assert)
    is std::terminate (lib/src/debug.hpp:52-62) and must not take the test
    process with it.  -> ("ok", parse dict) | ("refuses", message) | ("aborts", None)"""
    import pickle

    r, w = os.pipe()
    pid = os.fork()
    if pid == 0:
        os.close(r)
        try:
            import faulthandler

            faulthandler.disable()  # the abort of the child is an outcome, not a crash to report
            os.dup2(os.open(os.devnull, os.O_WRONLY), 2)
            h = RefProblem(lp_path, params, preprocess=True, affect=affect)
            want = h.instance()
            want["names"], want["bounds"] = h.variables()
            want["affected"] = h.affected()
            out = ("ok", want)
        except RuntimeError as e:
            out = ("refuses", str(e))
        with os.fdopen(w, "wb") as f:
            pickle.dump(out, f)
        os._exit(0)
    os.close(w)
    with os.fdopen(r, "rb") as f:
        data = f.read()
    os.waitpid(pid, 0)
    return pickle.loads(data) if data else ("aborts", None)
