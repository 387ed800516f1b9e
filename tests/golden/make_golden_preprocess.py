"""Golden outputs of the preprocessor: every tests/golden/lp_pp/*.lp through the
COMPILED REFERENCE (oracle/_ref/libbxref.so): baryonyx::preprocess, and for every
remaining variable of the first few baryonyx::affect(variable, true / false) on
the preprocessed problem (lib/src/preprocessor.cpp:857-903), then
make_merged_constraints.  Stored as <name>.json next to the LP text.  Needs
/root/reference (build container only):

    python tests/golden/make_golden_preprocess.py
"""
import glob
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import pyoracle as po  # noqa: E402

PRE_ORDERS = (0, 1, 4, 9)
MAX_AFFECT = 4


def plain(d):
    if d is None:
        return None
    out = {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in d.items()}
    out["bounds"] = [list(b) for b in out["bounds"]]
    out["affected"] = [list(a) for a in out["affected"]]
    return out


def one(lp, pre_order, affect=None):
    p = po.default_params()
    p.pre_order = pre_order
    status, d = po.ref_preprocess(lp, p, affect)
    return {"status": status, "problem": plain(d) if status == "ok" else None}


if __name__ == "__main__":
    for lp in sorted(glob.glob(os.path.join(HERE, "lp_pp", "*.lp"))):
        doc = {}
        for o in PRE_ORDERS:
            base = one(lp, o)
            doc[f"pre_order_{o}"] = base
            if base["status"] != "ok":
                continue
            for v in range(min(base["problem"]["n"], MAX_AFFECT)):
                for b in (1, 0):
                    doc[f"pre_order_{o}_affect_{v}_{b}"] = one(lp, o, [(v, b)])
        with open(lp[:-3] + ".json", "w") as f:
            json.dump(doc, f, separators=(",", ":"))
        d = doc["pre_order_0"]
        print(os.path.basename(lp), d["status"],
              d["problem"] and (d["problem"]["m"], d["problem"]["n"], len(d["problem"]["affected"])), len(doc), "cases")
