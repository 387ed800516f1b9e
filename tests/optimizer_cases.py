"""Cases of the optimizer-loop parity tests (thread = 1): shared by the CPU pin
(oracle vs compiled reference), the golden generator and the GPU tests."""
from __future__ import annotations

from baryonyx_b200 import synth

# name -> (instance factory, solver_parameters overrides, master seed, restarts)
CASES = {
    "spp_eq01": (lambda: synth.spp(m=40, n=300, lo=3, hi=6, seed=2, cost_hi=5),
                 dict(limit=300, pushes_limit=4, pushing_iteration_limit=6, init_population_size=20,
                      init_kappa_improve_stop=0.06), 42, 40),
    "spp_eq01_feasible": (lambda: synth.spp(m=20, n=120, lo=3, hi=5, seed=3, cost_hi=9),
                          dict(limit=400, pushes_limit=4, pushing_iteration_limit=6, init_population_size=20,
                               init_kappa_improve_stop=0.06, kappa_step=0.01), 5, 40),
    "scp_ineq01": (lambda: synth.scp(m=30, n=120),
                   dict(limit=60, pushes_limit=4, pushing_iteration_limit=6, init_population_size=20,
                        init_kappa_improve_stop=0.04), 42, 30),
    "mixed_eq101": (lambda: synth.small_mixed(m=12, n=30, seed=7, family="101", equality=True),
                    dict(limit=80, pushes_limit=4, pushing_iteration_limit=6, init_population_size=20,
                         init_kappa_improve_stop=0.06), 5, 30),
    "mixed_ineq101_max": (lambda: synth.small_mixed(m=12, n=30, seed=8, family="101", equality=False,
                                                    maximize=True),
                          dict(limit=80, pushes_limit=4, pushing_iteration_limit=6, init_population_size=20,
                               init_kappa_improve_stop=0.06), 5, 30),
    "telebus_ineqZ": (lambda: synth.telebus(m=40, n=150, seed=4),
                      dict(limit=80, pushes_limit=4, pushing_iteration_limit=6, init_population_size=20,
                           init_kappa_improve_stop=0.06, order=3), 5, 30),
    "telebus_ineqZ_b": (lambda: synth.telebus(m=25, n=80, seed=11),
                        dict(limit=80, pushes_limit=4, pushing_iteration_limit=6, init_population_size=20,
                             init_kappa_improve_stop=0.06), 5, 30),
    "telebus_ineqZ_c": (lambda: synth.telebus(m=60, n=200, seed=6),
                        dict(limit=80, pushes_limit=4, pushing_iteration_limit=6, init_population_size=20,
                             init_kappa_improve_stop=0.06, order=4), 5, 20),
    "scp_ineq01_infeasibility_order": (lambda: synth.scp(m=30, n=120, seed=4),
                                       dict(limit=60, pushes_limit=3, pushing_iteration_limit=5,
                                            init_population_size=10, init_kappa_improve_stop=0.04,
                                            order=3), 9, 30),
}

TRACE_KEYS = ("x", "rng", "kappa", "pi", "P", "pop_x", "pop_value", "pop_remaining", "pop_hash", "pop_loop",
              "sweep_round", "sweep_push", "sweep_kappa", "sweep_len")


def case_params(name, mod):
    """Unclamped solver_parameters of a case as `mod.Params` (oracle.pyoracle or capi)."""
    p = mod.default_params()
    for k, v in CASES[name][1].items():
        setattr(p, k, v)
    return p
