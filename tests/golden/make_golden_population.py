#!/usr/bin/env python
"""Generates tests/golden/population_*.json.gz from the UNMODIFIED reference
compiled here (oracle/_ref/libbxref.so): the trace of tests/test_population.py's
scripted sequence of best_solution_recorder / storage / local_context calls
(constructor, reinit, try_update, try_advance, crossover, mutation) on the
reference's own classes.  The fixtures travel to the GPU box, where
/root/reference does not exist.

    python tests/golden/make_golden_population.py    # needs oracle/_ref/libbxref.so
"""
from __future__ import annotations

import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import pyoracle as po  # noqa: E402
from tests import util  # noqa: E402
from tests.test_population import CASES, population_script  # noqa: E402

MASTERS = (42, 7)
STEPS, DATA_SEED = 40, 11


def plain(o):
    """JSON-able copy: arrays and 64-bit integers as decimal strings (exact)."""
    if isinstance(o, dict):
        return {k: plain(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [plain(v) for v in o]
    if isinstance(o, np.ndarray):
        return [str(int(v)) for v in o]
    if isinstance(o, (np.integer, int)):
        return str(int(o))
    if isinstance(o, float):
        return o.hex()
    return o


def main():
    if not po.ref_available():
        raise SystemExit("oracle/_ref/libbxref.so is missing: run `make -C oracle ref` first")
    for name, make, over in CASES:
        inst = make()
        p = util.params(**over)
        lp = util.lp_of(inst)
        rp = po.RefProblem(lp, p)
        os.unlink(lp)
        cp = rp.params()
        out = {"case": name, "steps": STEPS, "data_seed": DATA_SEED, "masters": {}}
        for master in MASTERS:
            pop = po.Population(rp, master, threads=2)
            out["masters"][str(master)] = plain(
                population_script(pop, rp.n, cp.kappa_min, cp.kappa_max, STEPS, DATA_SEED))
            pop.close()
        rp.close()
        path = os.path.join(HERE, f"population_{name}.json.gz")
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(json.dumps(out, separators=(",", ":")).encode())
        print(name, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
