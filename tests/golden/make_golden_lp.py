"""Golden parses for the LP reader: every tests/golden/lp/*.lp read by the
COMPILED REFERENCE (oracle/_ref/libbxref.so: make_problem, lib/src/parser.cpp:1065-1273,
+ make_merged_constraints, lib/src/merged-constraint.cpp:297-343), stored as
<name>.json next to the LP text.  Needs /root/reference (build container only):

    python tests/golden/make_golden_lp.py
"""
import glob
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import pyoracle as po  # noqa: E402


def parse_with_reference(path: str, pre_order: int) -> dict:
    p = po.default_params()
    p.pre_order = pre_order
    r = po.RefProblem(path, p)
    inst = r.instance()
    names, bounds = r.variables()
    r.close()
    out = {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in inst.items()}
    out["names"], out["bounds"] = names, [list(b) for b in bounds]
    return out


if __name__ == "__main__":
    for lp in sorted(glob.glob(os.path.join(HERE, "lp", "*.lp"))):
        doc = {f"pre_order_{o}": parse_with_reference(lp, o) for o in range(12)}
        with open(lp[:-3] + ".json", "w") as f:
            json.dump(doc, f, separators=(",", ":"))
        d = doc["pre_order_0"]
        print(os.path.basename(lp), "m", d["m"], "n", d["n"], "type", d["problem_type"])
