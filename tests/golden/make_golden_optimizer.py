"""Golden traces of the optimizer loop at thread = 1, generated from the
COMPILED REFERENCE (oracle/_ref/libbxref_fast.so: /root/reference built by
oracle/Makefile, contract checks compiled out) -- bxref_opt_trace runs the
reference's own optimize_functor on its own best_solution_recorder.

    python tests/golden/make_golden_optimizer.py

Writes tests/golden/optimizer_<case>.npz.  The GPU box has no /root/reference:
these files are what pins the CUDA optimizer there."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np

from oracle import pyoracle as po
from tests import util
from tests.optimizer_cases import CASES, TRACE_KEYS, case_params


def main():
    for name in sorted(CASES):
        factory, _, seed, rounds = CASES[name]
        inst = factory()
        p = case_params(name, po)
        path = util.lp_of(inst)
        pb = po.RefProblem(path, p, fast=True)
        os.unlink(path)
        tr = pb.optimize_trace(seed, rounds, p.init_population_size)
        pb.close()
        out = os.path.join(util.GOLDEN_DIR, f"optimizer_{name}.npz")
        np.savez_compressed(out, **{k: tr[k] for k in TRACE_KEYS})
        print(name, "restarts", rounds, "sweeps", len(tr["sweep_len"]), "push sweeps", int(tr["sweep_push"].sum()),
              "feasible members", int((tr["pop_remaining"][-1] == 0).sum()), os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
