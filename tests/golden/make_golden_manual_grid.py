"""Writes tests/golden/manual_grid.json from the COMPILED REFERENCE's
manual_course (oracle/_ref/libbxref.so: bxref_manual_grid, ref_harness/
harness_manual.cpp): the 5 x 5 tables and a sample of the 3125 points, as hex
floats, for boxes without /root/reference.  Run from the repository root:
python tests/golden/make_golden_manual_grid.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle as po  # noqa: E402
from test_manual_grid import BASES, NAMES  # noqa: E402

POINTS = [0, 1, 4, 5, 24, 25, 124, 125, 624, 625, 1562, 2499, 3000, 3123, 3124]

cases = []
for over in BASES:
    p = po.default_params()
    for k, v in over.items():
        setattr(p, k, v)
    r = po.ref_manual_grid(p)
    assert len(r["theta"]) == 3125
    case = {"base": over, "points": POINTS,
            "tables_hex": [float.hex(float(v)) for v in r["tables"].reshape(-1)]}
    for k in NAMES:
        case[k + "_hex"] = [float.hex(float(r[k][i])) for i in POINTS]
    cases.append(case)
with open(os.path.join(ROOT, "tests", "golden", "manual_grid.json"), "w") as f:
    json.dump({"source": "compiled reference, manual-optimizer.cpp manual_course(…, 5)", "cases": cases}, f, indent=1)
print("written", len(cases), "cases")
