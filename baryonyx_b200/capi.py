"""ctypes binding of the C ABI (include/bxg.h) -- the call path of the tests and
of bench.py.  Loads the in-tree ``libbaryonyx_b200.so`` and fails loudly when
it is missing: the engine has no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BXG_LIB") or os.path.join(HERE, "libbaryonyx_b200.so")


class Params(C.Structure):
    """bxg_params: POD mirror of baryonyx::solver_parameters
    (lib/include/baryonyx/core:294-462)."""

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

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


EXPORTS = [
    "bxg_version", "bxg_last_error", "bxg_device_count", "bxg_default_params", "bxg_clamp_params",
    "bxg_problem_create", "bxg_problem_destroy", "bxg_problem_info", "bxg_which_problem_type",
    "bxg_batch_create", "bxg_batch_destroy", "bxg_batch_info", "bxg_batch_set_tail_value",
    "bxg_batch_begin", "bxg_batch_get_state", "bxg_batch_set_state", "bxg_batch_sweep",
    "bxg_batch_objective", "bxg_batch_solve", "bxg_batch_best", "bxg_batch_argbest",
    "bxg_batch_counters", "bxg_batch_reset_counters", "bxg_batch_path_counters", "bxg_batch_last_ms", "bxg_batch_device_ms", "bxg_batch_launches",
    "bxg_batch_words32", "bxg_batch_begin_device", "bxg_batch_set_lane_params", "bxg_manual_grid",
    "bxg_population_create", "bxg_population_destroy", "bxg_population_size",
    "bxg_population_master_state", "bxg_population_generate_seeds", "bxg_population_member", "bxg_population_worker_seed",
    "bxg_population_worker_state", "bxg_population_reinit", "bxg_population_crossover",
    "bxg_population_mutation", "bxg_population_try_update", "bxg_population_try_advance",
    "bxg_population_score", "bxg_optimize_round", "bxg_optimize_reinit", "bxg_optimize_run",
    "bxg_optimize", "bxg_population_stats", "bxg_comm_unique_id", "bxg_comm_init", "bxg_comm_destroy",
    "bxg_comm_info", "bxg_exchange_best",
    "bxg_bit_array_hash",
    "bxg_lp_read", "bxg_lp_parse", "bxg_lp_destroy", "bxg_lp_info", "bxg_lp_rows", "bxg_lp_objective",
    "bxg_lp_variable", "bxg_lp_variables", "bxg_lp_names",
    "bxg_lp_preprocess", "bxg_lp_affect", "bxg_lp_split", "bxg_lp_affected_count", "bxg_lp_affected",
]

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")

_lib = None


class BxgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"bxg error {code}: {msg}")
        self.code = code


def lib():
    """The loaded C ABI.  Raises when the native library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("BXG_LIB", LIB_PATH)  # BXG_LIB: experiment builds of the same ABI
    if not os.path.exists(path):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m baryonyx_b200.build` "
            "(nvcc, sm_100a).  There is no CPU fallback."
        )
    L = C.CDLL(path)
    L.bxg_last_error.restype = C.c_char_p
    L.bxg_default_params.argtypes = [C.POINTER(Params)]
    L.bxg_default_params.restype = None
    L.bxg_clamp_params.argtypes = [C.POINTER(Params), C.POINTER(Params)]
    L.bxg_clamp_params.restype = None
    L.bxg_problem_create.argtypes = [
        C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p, _i32p, C.c_int, C.c_int, C.c_int, _i32p, _f64p,
        C.c_double, C.c_int, C.POINTER(C.c_void_p),
    ]
    L.bxg_problem_destroy.argtypes = [C.c_void_p]
    L.bxg_problem_destroy.restype = None
    L.bxg_problem_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 5
    L.bxg_which_problem_type.argtypes = [C.c_int, _i32p, C.c_int]
    L.bxg_batch_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(Params), C.POINTER(C.c_void_p)]
    L.bxg_batch_destroy.argtypes = [C.c_void_p]
    L.bxg_batch_destroy.restype = None
    L.bxg_batch_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.bxg_batch_set_tail_value.argtypes = [C.c_void_p, C.c_double]
    L.bxg_batch_begin.argtypes = [C.c_void_p, _u32p, C.c_int]
    L.bxg_batch_get_state.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 7
    L.bxg_batch_set_state.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
    L.bxg_batch_sweep.argtypes = [C.c_void_p, C.c_int, _f64p, C.c_double, C.c_void_p]
    L.bxg_batch_objective.argtypes = [C.c_void_p, _f64p]
    L.bxg_batch_solve.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    L.bxg_batch_best.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.bxg_batch_argbest.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.bxg_batch_counters.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                     C.POINTER(C.c_double)]
    L.bxg_batch_reset_counters.argtypes = [C.c_void_p]
    L.bxg_batch_path_counters.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                          C.POINTER(C.c_int64)]
    L.bxg_batch_last_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.bxg_batch_launches.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.bxg_batch_words32.argtypes = [C.c_void_p]
    L.bxg_batch_begin_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise BxgError(rc, lib().bxg_last_error().decode())


def default_params() -> Params:
    p = Params()
    lib().bxg_default_params(C.byref(p))
    return p


def clamp_params(p: Params) -> Params:
    q = Params()
    lib().bxg_clamp_params(C.byref(p), C.byref(q))
    return q


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def manual_grid(p: Params, first: int = 0, capacity: int = 3125):
    """bxg_manual_grid: points [first, first + capacity) of the reference's manual
    auto-tuner grid as the five arrays Batch.set_lane_params takes, + the 5 x 5 tables."""
    names = ("theta", "delta", "kappa_min", "kappa_step", "init_policy_random")
    arrs = [np.zeros(capacity, np.float64) for _ in names]
    tables = np.zeros((5, 5), np.float64)
    count = C.c_int64()
    f = lib().bxg_manual_grid
    f.argtypes = [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 5 + [C.POINTER(C.c_int64), C.c_void_p]
    check(f(C.byref(p), first, capacity, *[_ptr(a) for a in arrs], C.byref(count), _ptr(tables)))
    out = {k: a[:count.value] for k, a in zip(names, arrs)}
    out["tables"] = tables
    return out


class Problem:
    """bxg_problem: merged constraints (CSR) + objective, resident on one GPU."""

    def __init__(self, inst: dict, device: int = 0):
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        self.inst = inst
        self.h = C.c_void_p()
        obj_var = i32(inst["obj_var"]) if len(inst["obj_var"]) else np.zeros(1, np.int32)
        obj_fac = (np.ascontiguousarray(inst["obj_factor"], np.float64) if len(inst["obj_factor"])
                   else np.zeros(1, np.float64))
        check(lib().bxg_problem_create(
            int(inst["m"]), int(inst["n"]), i32(inst["row_ptr"]), i32(inst["col"]), i32(inst["factor"]),
            i32(inst["cmin"]), i32(inst["cmax"]), int(inst["problem_type"]), int(inst["maximize"]),
            len(inst["obj_var"]), obj_var, obj_fac, float(inst["cost_constant"]), device, C.byref(self.h),
        ))
        self.m, self.n = int(inst["m"]), int(inst["n"])
        self.nnz = int(inst["row_ptr"][-1])
        self.device = device

    def batch(self, n_lanes: int, params: Params) -> "Batch":
        return Batch(self, n_lanes, params)

    def close(self):
        if self.h:
            lib().bxg_problem_destroy(self.h)
            self.h = C.c_void_p()


class Batch:
    """bxg_batch: n_lanes independent Solver instances (P, pi, x, rng, order)."""

    def __init__(self, pb: Problem, n_lanes: int, params: Params):
        self.pb = pb
        self.n_lanes = n_lanes
        self.params = params
        self.h = C.c_void_p()
        check(lib().bxg_batch_create(pb.h, n_lanes, C.byref(params), C.byref(self.h)))
        self.words = pb.n // 64 + 1

    def info(self):
        nl, d, t, nb = C.c_int(), C.c_double(), C.c_double(), C.c_int64()
        check(lib().bxg_batch_info(self.h, C.byref(nl), C.byref(d), C.byref(t), C.byref(nb)))
        return {"n_lanes": nl.value, "delta": d.value, "theta": t.value, "device_bytes": nb.value}

    def set_tail_value(self, v: float):
        check(lib().bxg_batch_set_tail_value(self.h, float(v)))

    def begin(self, seeds, reset_solver: bool = True):
        seeds = np.ascontiguousarray(seeds, np.uint32)
        assert len(seeds) == self.n_lanes
        check(lib().bxg_batch_begin(self.h, seeds, int(reset_solver)))

    def begin_device(self, d_x0_ptr: int, d_rng_ptr: int, reset_solver: bool = True):
        check(lib().bxg_batch_begin_device(self.h, C.c_void_p(d_x0_ptr), C.c_void_p(d_rng_ptr),
                                           int(reset_solver)))

    def words32(self) -> int:
        return lib().bxg_batch_words32(self.h)

    def state(self, lane: int = 0):
        pb = self.pb
        x = np.zeros(self.words, np.uint64)
        pi = np.zeros(pb.m, np.float64)
        P = np.zeros(max(pb.nnz, 1), np.float64)
        rng = C.c_uint32()
        ol = np.zeros(2 * max(pb.m, pb.n) + 2, np.int32)
        on = C.c_int()
        check(lib().bxg_batch_get_state(self.h, lane, _ptr(x), _ptr(pi), _ptr(P),
                                        C.cast(C.byref(rng), C.c_void_p), _ptr(ol),
                                        C.cast(C.byref(on), C.c_void_p), None))
        return {"x": x, "pi": pi, "P": P[: pb.nnz], "rng": rng.value, "order": ol, "order_len": on.value}

    def norm_cost(self):
        c = np.zeros(self.pb.n, np.float64)
        check(lib().bxg_batch_get_state(self.h, 0, None, None, None, None, None, None, _ptr(c)))
        return c

    def set_state(self, lane, x_words=None, pi=None, P=None, rng=None):
        xw = None if x_words is None else np.ascontiguousarray(x_words, np.uint64)
        pi_ = None if pi is None else np.ascontiguousarray(pi, np.float64)
        P_ = None if P is None else np.ascontiguousarray(P, np.float64)
        r = None if rng is None else C.c_uint32(rng)
        check(lib().bxg_batch_set_state(self.h, lane, _ptr(xw), _ptr(pi_), _ptr(P_),
                                        None if r is None else C.cast(C.byref(r), C.c_void_p)))

    def sweep(self, kappa, push: int = 0, obj_amp: float = 0.0):
        k = np.ascontiguousarray(np.broadcast_to(np.asarray(kappa, np.float64), (self.n_lanes,)))
        rem = np.zeros(self.n_lanes, np.int32)
        check(lib().bxg_batch_sweep(self.h, int(push), k, float(obj_amp), _ptr(rem)))
        return rem

    def objective(self):
        v = np.zeros(self.n_lanes, np.float64)
        check(lib().bxg_batch_objective(self.h, v))
        return v

    def solve(self):
        st = np.zeros(self.n_lanes, np.int32)
        val = np.zeros(self.n_lanes, np.float64)
        rem = np.zeros(self.n_lanes, np.int32)
        loop = np.zeros(self.n_lanes, np.int64)
        check(lib().bxg_batch_solve(self.h, _ptr(st), _ptr(val), _ptr(rem), _ptr(loop)))
        return {"status": st, "value": val, "remaining": rem, "loop": loop}

    def best(self, lane: int):
        x = np.zeros(self.pb.n, np.int8)
        check(lib().bxg_batch_best(self.h, lane, _ptr(x)))
        return x

    def argbest(self):
        l, v, r = C.c_int(), C.c_double(), C.c_int()
        check(lib().bxg_batch_argbest(self.h, C.byref(l), C.byref(v), C.byref(r)))
        return l.value, v.value, r.value

    def counters(self):
        u, s, b = C.c_int64(), C.c_int64(), C.c_double()
        check(lib().bxg_batch_counters(self.h, C.byref(u), C.byref(s), C.byref(b)))
        return {"updates": u.value, "sweeps": s.value, "algorithmic_bytes": b.value}

    def path_counters(self):
        """Updates and cycles by selection path: exact, sorted, two smallest, sign boundary."""
        u = (C.c_int64 * 4)()
        c = (C.c_int64 * 4)()
        ph = (C.c_int64 * 4)()
        check(lib().bxg_batch_path_counters(self.h, u, c, ph))
        return {"updates": list(u), "cycles": list(c), "phase_cycles": list(ph)}

    def set_lane_params(self, theta=None, delta=None, kappa_min=None, kappa_step=None,
                        init_policy_random=None):
        """Per-lane theta / delta / kappa_min / kappa_step / init_policy_random (parameter sweeps)."""
        arrs = []
        for a in (theta, delta, kappa_min, kappa_step, init_policy_random):
            if a is None:
                arrs.append(None)
            else:
                a = np.ascontiguousarray(a, np.float64)
                assert a.shape == (self.n_lanes,)
                arrs.append(a)
        f = lib().bxg_batch_set_lane_params
        f.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        check(f(self.h, *[_ptr(a) for a in arrs]))

    def reset_counters(self):
        check(lib().bxg_batch_reset_counters(self.h))

    def last_ms(self) -> float:
        ms = C.c_float()
        check(lib().bxg_batch_last_ms(self.h, C.byref(ms)))
        return ms.value

    def device_ms(self) -> float:
        """Device time of all timed launches since reset_counters (CUDA events)."""
        ms = C.c_double()
        lib().bxg_batch_device_ms.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        check(lib().bxg_batch_device_ms(self.h, C.byref(ms)))
        return ms.value

    def launches(self) -> int:
        n = C.c_int64()
        check(lib().bxg_batch_launches(self.h, C.byref(n)))
        return n.value

    def close(self):
        if self.h:
            lib().bxg_batch_destroy(self.h)
            self.h = C.c_void_p()


_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")


def _bind_population(L):
    if getattr(L, "_pop_bound", False):
        return
    L.bxg_population_create.argtypes = [C.c_void_p, C.POINTER(Params), C.c_uint64, C.c_int,
                                        C.POINTER(C.c_void_p)]
    L.bxg_population_destroy.argtypes = [C.c_void_p]
    L.bxg_population_size.argtypes = [C.c_void_p]
    L.bxg_population_master_state.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    L.bxg_population_generate_seeds.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.bxg_population_member.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_double),
                                        C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]
    L.bxg_population_worker_seed.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
    L.bxg_population_worker_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.bxg_population_reinit.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, _u64p,
                                        C.POINTER(C.c_double)]
    L.bxg_population_crossover.argtypes = [C.c_void_p, C.c_int, _u64p]
    L.bxg_population_mutation.argtypes = [C.c_void_p, C.c_int, _u64p]
    L.bxg_population_try_update.argtypes = [C.c_void_p, C.c_int, _u64p, C.c_double, C.c_int64]
    L.bxg_population_try_advance.argtypes = [C.c_void_p, C.c_int, _u64p, C.c_int, C.c_int64]
    L.bxg_population_score.argtypes = [C.c_void_p, _u64p, C.c_int, _i32p, _f64p]
    L.bxg_optimize_round.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.bxg_optimize.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int64, C.c_int, C.c_void_p,
                               C.c_void_p, C.POINTER(C.c_int64)]
    L.bxg_population_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    L.bxg_comm_unique_id.argtypes = [C.c_void_p]
    L.bxg_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.bxg_comm_destroy.argtypes = [C.c_void_p]
    L.bxg_comm_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64)]
    L.bxg_exchange_best.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    L.bxg_optimize_reinit.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.bxg_optimize_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.bxg_bit_array_hash.argtypes = [C.c_int, _u64p]
    L.bxg_bit_array_hash.restype = C.c_uint64
    L._pop_bound = True


UPDATE_FN = C.CFUNCTYPE(None, C.c_int, C.c_double, C.c_int64, C.c_double, C.c_int64, C.c_void_p)


class Comm:
    """bxg_comm: the NCCL communicator of the incumbent exchange (one per rank)."""

    @staticmethod
    def unique_id() -> bytes:
        _bind_population(lib())
        buf = C.create_string_buffer(128)
        check(lib().bxg_comm_unique_id(buf))
        return buf.raw

    def __init__(self, unique_id: bytes, rank: int, nranks: int, device: int):
        _bind_population(lib())
        self.h = C.c_void_p()
        check(lib().bxg_comm_init(unique_id, rank, nranks, device, C.byref(self.h)))
        self.rank, self.nranks = rank, nranks

    def exchanges(self) -> int:
        n = C.c_int64()
        check(lib().bxg_comm_info(self.h, None, None, C.byref(n)))
        return n.value

    def close(self):
        if self.h:
            lib().bxg_comm_destroy(self.h)
            self.h = C.c_void_p()


class Population:
    """bxg_population: the optimizer's candidate vectors and per-worker random
    streams (best_solution_recorder / local_context of the reference)."""

    def __init__(self, pb: Problem, params: Params, master_seed: int, workers: int = 1):
        _bind_population(lib())
        self.pb = pb
        self.n = pb.n
        self.words = pb.n // 64 + 1
        self.h = C.c_void_p()
        check(lib().bxg_population_create(pb.h, C.byref(params), master_seed, workers, C.byref(self.h)))
        st = C.c_uint32()
        check(lib().bxg_population_master_state(self.h, C.byref(st)))
        self.rng_after_ctor = st.value

    def size(self) -> int:
        return lib().bxg_population_size(self.h)

    def member(self, i: int):
        x = np.zeros(self.words, np.uint64)
        v, r, h, l = C.c_double(), C.c_int(), C.c_uint64(), C.c_int64()
        check(lib().bxg_population_member(self.h, i, x.ctypes.data_as(C.c_void_p), C.byref(v), C.byref(r),
                                          C.byref(h), C.byref(l)))
        return {"x": x, "value": v.value, "remaining": r.value, "hash": h.value, "loop": l.value}

    def members(self):
        return [self.member(i) for i in range(self.size())]

    def generate_seeds(self, count: int) -> np.ndarray:
        """generate_seed on the master engine (after the constructor)."""
        out = np.zeros(count, np.uint64)
        check(lib().bxg_population_generate_seeds(self.h, count, out.ctypes.data_as(C.c_void_p)))
        return out

    def ctx(self, worker: int, seed: int):
        check(lib().bxg_population_worker_seed(self.h, worker, seed))

    def ctx_rng(self, worker: int) -> int:
        st = C.c_uint32()
        check(lib().bxg_population_worker_state(self.h, worker, C.cast(C.byref(st), C.c_void_p), None))
        return st.value

    def set_ctx_rng(self, worker: int, state: int):
        st = C.c_uint32(state)
        check(lib().bxg_population_worker_state(self.h, worker, None, C.cast(C.byref(st), C.c_void_p)))

    def reinit(self, worker: int, kappa_min: float, kappa_max: float, x: np.ndarray) -> float:
        k = C.c_double()
        check(lib().bxg_population_reinit(self.h, worker, kappa_min, kappa_max, x, C.byref(k)))
        return k.value

    def crossover(self, worker: int, x: np.ndarray):
        check(lib().bxg_population_crossover(self.h, worker, x))

    def mutation(self, worker: int, x: np.ndarray):
        check(lib().bxg_population_mutation(self.h, worker, x))

    def try_update(self, worker: int, x: np.ndarray, value: float, loop: int):
        check(lib().bxg_population_try_update(self.h, worker, x, value, loop))

    def try_advance(self, worker: int, x: np.ndarray, remaining: int, loop: int):
        check(lib().bxg_population_try_advance(self.h, worker, x, remaining, loop))

    def score(self, xs: np.ndarray):
        xs = np.ascontiguousarray(xs, np.uint64).reshape(-1, self.words)
        rem = np.zeros(len(xs), np.int32)
        val = np.zeros(len(xs), np.float64)
        check(lib().bxg_population_score(self.h, xs.reshape(-1), len(xs), rem, val))
        return rem, val

    def optimize_round(self, batch: "Batch", first_round: bool):
        """One optimizer call on every lane of `batch` (lane l = worker l)."""
        st = np.zeros(batch.n_lanes, np.int32)
        val = np.zeros(batch.n_lanes, np.float64)
        rem = np.zeros(batch.n_lanes, np.int32)
        check(lib().bxg_optimize_round(batch.h, self.h, int(first_round), _ptr(st), _ptr(val), _ptr(rem)))
        return {"status": st, "value": val, "remaining": rem}

    def optimize_reinit(self, batch: "Batch", first_round: bool):
        """First half of a round (parity tap): reinit of every lane; returns the kappas."""
        k = np.zeros(batch.n_lanes, np.float64)
        check(lib().bxg_optimize_reinit(batch.h, self.h, int(first_round), _ptr(k)))
        return k

    def optimize_run(self, batch: "Batch"):
        """Second half of a round: the lanes' runs with the offers to the population."""
        st = np.zeros(batch.n_lanes, np.int32)
        val = np.zeros(batch.n_lanes, np.float64)
        rem = np.zeros(batch.n_lanes, np.int32)
        check(lib().bxg_optimize_run(batch.h, self.h, _ptr(st), _ptr(val), _ptr(rem)))
        return {"status": st, "value": val, "remaining": rem}

    def optimize(self, batch: "Batch", time_limit: float = 0.0, max_rounds: int = 0, comm: "Comm | None" = None,
                 exchange_every: int = 0, update=None) -> int:
        """bxg_optimize: rounds until the time limit or `max_rounds`; `update(remaining, value,
        loop, duration, restarts)` about once a second.  Returns the rounds done."""
        done = C.c_int64()
        cb = None
        if update is not None:
            cb = UPDATE_FN(lambda rem, val, loop, dur, restarts, user: update(rem, val, loop, dur, restarts))
        check(lib().bxg_optimize(batch.h, self.h, comm.h if comm else None, float(time_limit), int(max_rounds),
                                 int(exchange_every), C.cast(cb, C.c_void_p) if cb else None, None,
                                 C.byref(done)))
        return done.value

    def stats(self):
        r, t = C.c_int64(), C.c_double()
        check(lib().bxg_population_stats(self.h, C.byref(r), C.byref(t)))
        return {"rounds": r.value, "first_solution_s": None if t.value < 0 else t.value}

    def exchange(self, batch: "Batch", comm: "Comm", wait: bool = True):
        w = C.c_int(-1)
        check(lib().bxg_exchange_best(batch.h, self.h, comm.h, C.byref(w) if wait else None))
        return w.value

    def hash(self, x: np.ndarray) -> int:
        return lib().bxg_bit_array_hash(self.n, x)

    def close(self):
        if self.h:
            lib().bxg_population_destroy(self.h)
            self.h = C.c_void_p()


def _lp_types(L):
    L.bxg_lp_read.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    L.bxg_lp_parse.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]
    L.bxg_lp_destroy.argtypes = [C.c_void_p]
    L.bxg_lp_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 5 + [C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.bxg_lp_rows.argtypes = [C.c_void_p, _i32p, _i32p, _i32p, _i32p, _i32p]
    L.bxg_lp_objective.argtypes = [C.c_void_p, _i32p, _f64p]
    L.bxg_lp_variable.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 3
    L.bxg_lp_variable.restype = C.c_char_p
    L.bxg_lp_variables.argtypes = [C.c_void_p, _i32p, _i32p, _i32p]
    L.bxg_lp_names.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.bxg_lp_preprocess.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    L.bxg_lp_affect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.bxg_lp_split.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.bxg_lp_affected_count.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.bxg_lp_affected.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    L.bxg_lp_affected.restype = C.c_char_p


def _lp_export(L, h) -> dict:
    v = [C.c_int() for _ in range(5)]
    cc, nobj = C.c_double(), C.c_int()
    check(L.bxg_lp_info(h, *[C.byref(a) for a in v], C.byref(cc), C.byref(nobj)))
    m, n, nnz, ptype, maximize = [a.value for a in v]
    row_ptr = np.zeros(m + 1, np.int32)
    col, fac = np.zeros(max(nnz, 1), np.int32), np.zeros(max(nnz, 1), np.int32)
    cmin, cmax = np.zeros(max(m, 1), np.int32), np.zeros(max(m, 1), np.int32)
    check(L.bxg_lp_rows(h, row_ptr, col, fac, cmin, cmax))
    ov, of = np.zeros(max(nobj.value, 1), np.int32), np.zeros(max(nobj.value, 1), np.float64)
    check(L.bxg_lp_objective(h, ov, of))
    lo, hi, ty = (np.zeros(max(n, 1), np.int32) for _ in range(3))
    check(L.bxg_lp_variables(h, lo, hi, ty))
    need = C.c_size_t()
    check(L.bxg_lp_names(h, None, 0, C.byref(need)))
    raw = C.create_string_buffer(max(need.value, 1))
    check(L.bxg_lp_names(h, raw, need.value, C.byref(need)))
    names = raw.raw[: need.value].decode().split("\0")[:n]
    bounds = list(zip(lo[:n].tolist(), hi[:n].tolist(), ty[:n].tolist()))
    na = C.c_int()
    check(L.bxg_lp_affected_count(h, C.byref(na)))
    affected = []
    for i in range(na.value):
        val = C.c_int()
        affected.append((L.bxg_lp_affected(h, i, C.byref(val)).decode(), val.value))
    return {"m": m, "n": n, "row_ptr": row_ptr, "col": col[:nnz], "factor": fac[:nnz], "cmin": cmin[:m],
            "cmax": cmax[:m], "problem_type": ptype, "maximize": maximize, "obj_var": ov[: nobj.value],
            "obj_factor": of[: nobj.value], "cost_constant": cc.value, "names": names, "bounds": bounds,
            "affected": affected}


def read_lp(path: str | None = None, text: str | None = None, pre_order: int = 0, preprocess: bool = False,
            affect: list | None = None) -> dict:
    """LP file (or text) -> the instance dict Problem takes: the engine's own
    reader + constraint merging (bxg_lp_*), no device needed.  preprocess: run
    bxg_lp_preprocess on the parse (the reference's preprocessor "all");
    affect: [(variable, value), ...] applied one after the other with
    bxg_lp_affect on the result."""
    L = lib()
    _lp_types(L)
    h = C.c_void_p()
    # BXG_LP_RAW (-1) in front of the preprocessor: the reference merges after it
    first = -1 if preprocess else pre_order
    if path is not None:
        check(L.bxg_lp_read(path.encode(), first, C.byref(h)))
    else:
        raw = text.encode()
        check(L.bxg_lp_parse(raw, len(raw), first, C.byref(h)))
    try:
        steps = ([("pp",)] if preprocess else []) + [("af", int(v), int(b)) for v, b in (affect or [])]
        for i, st in enumerate(steps):
            h2 = C.c_void_p()
            order = pre_order if i + 1 == len(steps) else -1  # rows are merged once, at the end
            if st[0] == "pp":
                check(L.bxg_lp_preprocess(h, order, C.byref(h2)))
            else:
                check(L.bxg_lp_affect(h, st[1], st[2], order, C.byref(h2)))
            L.bxg_lp_destroy(h)
            h = h2
        return _lp_export(L, h)
    finally:
        L.bxg_lp_destroy(h)


def split_lp(path: str, variable: int, pre_order: int = 0, preprocess: bool = True) -> tuple:
    """bxg_lp_split on the (preprocessed) parse: the two halves as instance dicts."""
    L = lib()
    _lp_types(L)
    h = C.c_void_p()
    check(L.bxg_lp_read(path.encode(), -1 if preprocess else pre_order, C.byref(h)))
    hs = [h]
    try:
        if preprocess:
            h2 = C.c_void_p()
            check(L.bxg_lp_preprocess(h, -1, C.byref(h2)))
            hs.append(h2)
        a, b = C.c_void_p(), C.c_void_p()
        check(L.bxg_lp_split(hs[-1], variable, pre_order, C.byref(a), C.byref(b)))
        hs += [a, b]
        return _lp_export(L, a), _lp_export(L, b)
    finally:
        for x in hs:
            L.bxg_lp_destroy(x)
