"""time bxg_lp_read against the compiled reference's
make_problem + make_merged_constraints (oracle/_ref) on the same LP text.
Host code on one core each; build container only (needs oracle/_ref).
Test infrastructure (it runs the checker next to the product), hence under tests/;
not collected by pytest.

    python tests/lp_reader_timing.py [m n]      # rail-shaped instance, default 2000 x 200000
    python tests/lp_reader_timing.py m n --no-reference   # C5 scale: the reference's quadratic
                                                # look-ups make it impractical there; the parse is
                                                # checked against the instance the file was written from
"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from baryonyx_b200 import capi, synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def timed(f, reps=3):
    best = float("inf")
    for _ in range(reps):
        t = time.perf_counter()
        out = f()
        best = min(best, time.perf_counter() - t)
    return best, out


def main():
    no_ref = "--no-reference" in sys.argv
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    m, n = (int(args[0]), int(args[1])) if len(args) == 2 else (2000, 200000)
    inst = synth.rail(m=m, n=n)
    path = synth.write_lp(inst, os.path.join(tempfile.mkdtemp(), "rail.lp"))
    size = os.path.getsize(path)
    import ctypes as C

    def mine():
        h = C.c_void_p()
        capi.check(capi.lib().bxg_lp_read(path.encode(), 0, C.byref(h)))
        capi.lib().bxg_lp_destroy.argtypes = [C.c_void_p]
        capi.lib().bxg_lp_destroy(h)

    def ref():
        r = po.RefProblem(path, po.default_params())
        r.close()

    if no_ref:
        t_mine, _ = timed(mine, reps=2)
        L = capi.lib()
        h = C.c_void_p()
        capi.check(L.bxg_lp_read(path.encode(), 0, C.byref(h)))
        v = [C.c_int() for _ in range(5)]
        cc, nobj = C.c_double(), C.c_int()
        L.bxg_lp_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 5 + [C.POINTER(C.c_double), C.POINTER(C.c_int)]
        capi.check(L.bxg_lp_info(h, *[C.byref(a) for a in v], C.byref(cc), C.byref(nobj)))
        mm, nn, nnz = v[0].value, v[1].value, v[2].value
        row_ptr, col, fac = np.zeros(mm + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz, np.int32)
        cmin, cmax = np.zeros(mm, np.int32), np.zeros(mm, np.int32)
        i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
        L.bxg_lp_rows.argtypes = [C.c_void_p] + [i32p] * 5
        capi.check(L.bxg_lp_rows(h, row_ptr, col, fac, cmin, cmax))
        L.bxg_lp_destroy.argtypes = [C.c_void_p]
        L.bxg_lp_destroy(h)
        assert (mm, nn) == (inst["m"], inst["n"])
        for k, a in (("row_ptr", row_ptr), ("col", col), ("factor", fac), ("cmin", cmin), ("cmax", cmax)):
            assert np.array_equal(a, np.asarray(inst[k])), k
        print({"file_MB": round(size / 1e6, 1), "m": mm, "n": nn, "nnz": nnz, "bxg_lp_read_s": round(t_mine, 3),
               "MB_per_s": round(size / 1e6 / t_mine, 1), "reference_s": None, "checked_against": "written instance"})
        return
    d = capi.read_lp(path=path)
    r = po.RefProblem(path, po.default_params())
    want = r.instance()
    r.close()
    for k in ("row_ptr", "col", "factor", "cmin", "cmax", "obj_var", "obj_factor"):
        assert np.array_equal(d[k], want[k]), k
    t_mine, _ = timed(mine)
    t_ref, _ = timed(ref)
    print({"file_MB": round(size / 1e6, 1), "m": d["m"], "n": d["n"], "nnz": int(d["row_ptr"][-1]),
           "bxg_lp_read_s": round(t_mine, 3), "reference_s": round(t_ref, 3),
           "MB_per_s": round(size / 1e6 / t_mine, 1), "ratio": round(t_ref / t_mine, 2)})


if __name__ == "__main__":
    main()
