#!/usr/bin/env python
"""bench.py -- constraint-updates/s of the B200 sweep engine (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path

Default (--mode optimize): the reference's OPTIMIZER on the workload, its
workers batched as lanes.  A *step* is `--rounds` calls of every worker's loop
body (optimize_functor::operator(), lib/src/itm-optimizer-common.hpp:658-750:
reinit by crossover / mutation from the shared population, the sweeps with the
kappa schedule, the push phase, every solution offered to the population) --
bxg_optimize through the C ABI (include/bxg.h), population and lanes resident
in HBM.  Workload at N=1 = BASELINE.json configs[1]: the SPP-shaped
500 x 10 000 set-partitioning instance, 1024 workers.  For N>1 every rank runs
its own `lanes` workers on its own copy of the population (weak scaling) and
the incumbent exchange -- one ncclAllGather of the ranks' best members on the
batch's stream, bxg_exchange_best -- runs INSIDE the timed region every
`--exchange-every` rounds.  The reference arm runs the reference's own
optimize_problem on all host threads.

--mode solve is round 1's workload: `lanes` independent restarts of
solver_functor (lib/src/itm-solver-common.hpp:135-225) for `limit` sweeps,
bxg_batch_begin + bxg_batch_solve.

The JSON line carries the device-resident number (`value`), the end-to-end
number through the C ABI with host buffers (`e2e`), the roofline of the
dominant kernel and the reference's CPU implementation timed on this box's
host cores (`cpu_baseline`, rank 0, N=1 only).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "constraint-updates/s"
UNIT = "updates/s"
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="spp", choices=["spp", "scp", "csplib022", "telebus", "rail-small", "rail"])
    ap.add_argument("--lanes", type=int, default=1024, help="restarts per GPU per step")
    ap.add_argument("--mode", default="optimize", choices=["optimize", "solve"])
    ap.add_argument("--limit", type=int, default=0,
                    help="sweeps per restart (solver_parameters.limit); 0 = the reference's default 1000 in "
                         "optimize mode, 100 in solve mode")
    ap.add_argument("--rounds", type=int, default=1, help="optimizer rounds per step (optimize mode)")
    ap.add_argument("--exchange-every", type=int, default=1, help="incumbent exchange every this many rounds (N>1)")
    ap.add_argument("--seed", type=int, default=42, help="solver_parameters.seed of the optimizer")
    ap.add_argument("--cpu-seconds", type=float, default=0.0, help="time_limit of one reference optimizer run")
    ap.add_argument("--cpu-builds", default="both", choices=["both", "o2", "o3"],
                    help="which builds of the reference to time (the faster is quoted)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-lanes", type=int, default=0, help="restarts in the CPU sample (0 = 2 x cores)")
    args = ap.parse_args()
    if args.limit <= 0:
        args.limit = 1000 if args.mode == "optimize" else 100
    return args


def make_instance(name):
    from baryonyx_b200 import synth

    if name == "spp":
        return synth.spp(), "SPP-shaped 500x10000 set partitioning (BASELINE configs[1])"
    if name == "scp":
        return synth.scp(), "scp41-shaped 200x1000 set covering (BASELINE configs[0])"
    if name == "csplib022":
        return synth.csplib022(), "csplib022-shaped 30x50000, 101 solver (BASELINE configs[2])"
    if name == "telebus":
        return synth.telebus(), "telebus-shaped 1500x7000, Z solver (BASELINE configs[3])"
    if name == "rail":
        return synth.rail(), "rail4284-shaped 4284x1092610 (BASELINE configs[4])"
    if name == "rail-small":
        return synth.rail(m=1071, n=273152), "rail4284-shaped at 1/4 scale"
    raise ValueError(name)


def bench_params(args, mod):
    """solver_parameters of the bench: the reference's defaults except `limit`."""
    p = mod.default_params()
    p.limit = args.limit
    if args.workload == "telebus":
        p.order = 3  # infeasibility_decr: BASELINE configs[3] names the infeasibility ordering
    return mod.clamp_params(p)


def measured_traffic(workload, lanes, limit, mode="solve"):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_solve launch from
    the committed `ncu --set full` capture of this same workload
    (profiles/traffic.json, written from the .ncu-rep by profiles/ncu_summary.sh);
    None when the capture was taken on another configuration."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            d = json.load(f)
        if (d.get("workload"), d.get("lanes"), d.get("limit"), d.get("mode", "solve")) == (workload, lanes, limit, mode):
            return float(d["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        for k in ("hbm_gbs", "hbm_gb_s", "hbm_GBs"):
            if k in d:
                return float(d[k]), "measured"
    except Exception:
        pass
    return HBM_FALLBACK_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="bx_clocks_", suffix=".csv")
            self.f = os.fdopen(fd, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            with open(self.path) as f:
                for line in f:
                    c = [t.strip() for t in line.split(",")]
                    if len(c) < 8:
                        continue
                    try:
                        sm.append(float(c[1]))
                        mx.append(float(c[2]))
                    except ValueError:
                        continue
                    for nm, v in zip(names, c[4:8]):
                        if v.lower().startswith("active"):
                            reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference (oracle/_ref) when its
# prebuilt library is present, else the C restatement (oracle/liboracle.so)
# ---------------------------------------------------------------------------
def cpu_solve_sample(inst, args, n_lanes, seed0):
    """Runs `n_lanes` restarts of the bench workload on the host cores with the
    reference's own solve_problem<> (counting Solver subclass), one restart
    per worker thread at a time.  Returns (updates, sweeps, seconds, kind,
    threads)."""
    from baryonyx_b200 import synth
    from oracle import pyoracle as po

    cores = os.cpu_count() or 1
    threads = min(cores, n_lanes)
    seeds = [seed0 + i for i in range(n_lanes)]
    results = [None] * n_lanes

    if po.ref_available(fast=True):
        kind = "reference"
        p = po.default_params()
        p.limit = args.limit
        if args.workload == "telebus":
            p.order = 3
        fd, path = tempfile.mkstemp(suffix=".lp", prefix="bx_bench_")
        os.close(fd)
        synth.write_lp(inst, path)
        pb = po.RefProblem(path, p, fast=True)
        os.unlink(path)
        solve = lambda s: pb.solve(s)
    else:
        kind = "port"
        po.build_oracle()
        cp = bench_params(args, po)
        pb = po.OracleProblem(inst)
        solve = lambda s: pb.solve(cp, s)

    nxt = [0]
    lock = threading.Lock()

    def worker():
        while True:
            with lock:
                i = nxt[0]
                nxt[0] += 1
            if i >= n_lanes:
                return
            results[i] = solve(seeds[i])

    t0 = time.perf_counter()
    ts = [threading.Thread(target=worker) for _ in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    pb.close()
    upd = sum(r["updates"] for r in results)
    swp = sum(r["sweeps"] for r in results)
    return upd, swp, dt, kind, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    inst, wname = make_instance(args.workload)
    cores = os.cpu_count() or 1
    # a bounded sample of the workload per step: a quarter of the restarts
    # (about 3 s on 16 cores), so that K steps end within a minute or two
    n_lanes = args.cpu_lanes or max(2 * cores, args.lanes // 4)
    for _ in range(args.warmup):
        cpu_solve_sample(inst, args, min(n_lanes, cores), 1)
    upd = 0
    sec = 0.0
    kind = threads = None
    for s in range(args.steps):
        u, _, dt, kind, threads = cpu_solve_sample(inst, args, n_lanes, 1 + s * n_lanes)
        upd += u
        sec += dt
    value = upd / sec
    sample = f"{n_lanes} of {args.lanes} restarts per step, limit={args.limit}, {threads} host threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_of(args, wname, inst),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------
# the engine
# ---------------------------------------------------------------------------
def pinned_like(a):
    """Copy of a numpy array in pinned host memory (a numpy view on a torch
    pinned tensor; the tensor is kept alive by the returned object's base)."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(a)).clone().pin_memory()
    v = t.numpy()
    pinned_like.keep.append(t)
    return v


pinned_like.keep = []


def run_engine(args):
    import torch
    import torch.distributed as dist

    from baryonyx_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    inst, wname = make_instance(args.workload)
    cp = bench_params(args, capi)
    L = args.lanes
    nnz = int(inst["row_ptr"][-1])

    pb = capi.Problem(inst, device=local)
    bt = pb.batch(L, cp)
    info = bt.info()

    def seeds_of(step):
        base = 1 + (rank * 1_000_003 + step) * L
        return (np.arange(L, dtype=np.uint64) + base).astype(np.uint32)

    def one_step(step):
        bt.begin(seeds_of(step), reset_solver=True)
        ms = bt.last_ms()
        out = bt.solve()
        return ms, bt.last_ms(), out

    for s in range(args.warmup):
        one_step(s)

    bt.reset_counters()
    launches0 = bt.launches()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0
    solve_ms = 0.0
    for s in range(args.steps):
        a, b, out = one_step(args.warmup + s)
        dev_ms += a + b
        solve_ms += b
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    cnt = bt.counters()
    launches = bt.launches() - launches0
    lane, best_value, best_rem = bt.argbest()

    # max over ranks of the step time, sum over ranks of the work
    t = torch.tensor([wall, dev_ms], dtype=torch.float64, device="cuda")
    w = torch.tensor([float(cnt["updates"]), float(cnt["sweeps"]), cnt["algorithmic_bytes"],
                      float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    # incumbent exchange: who holds the best solution over the
    # ranks (all_reduce MIN over NCCL), and its bits (broadcast from the winner)
    from baryonyx_b200 import exchange

    xbits = bt.best(lane).astype(np.uint8)
    xw = np.zeros(inst["n"] // 64 + 1, np.uint64)
    packed = np.packbits(xbits, bitorder="little")
    xw.view(np.uint8)[: len(packed)] = packed
    winner, inc_rem, inc_value, inc_x = exchange.exchange_best(
        best_rem, best_value, xw, maximize=bool(inst["maximize"]))
    wall_max, dev_ms_max = t.tolist()
    updates, sweeps, alg_bytes, launches_all = w.tolist()

    # ---- end to end through the C ABI with host buffers ---------------------
    e2e = None
    if not args.no_e2e:
        keys = ("row_ptr", "col", "factor", "cmin", "cmax", "obj_var", "obj_factor")
        hinst = dict(inst)
        for k in keys:
            hinst[k] = pinned_like(inst[k])
        h2d = sum(hinst[k].nbytes for k in keys) + 4 * L
        d2h = 0
        step_seeds = [pinned_like(seeds_of(args.warmup + s)) for s in range(args.steps)]
        split = [0.0] * 6  # ms in: problem create, batch create, begin, solve, results, destroy

        def e2e_step(seeds):
            """What a caller of the C ABI does for one batch: merged constraints
            and seeds from host memory to a new problem + batch, the run, the
            best vector and the per-lane results back to the host."""
            t0 = time.perf_counter()
            pb2 = capi.Problem(hinst, device=local)
            t1 = time.perf_counter()
            bt2 = pb2.batch(L, cp)
            t2 = time.perf_counter()
            bt2.begin(seeds, reset_solver=True)
            t3 = time.perf_counter()
            out2 = bt2.solve()
            t4 = time.perf_counter()
            ln, _, _ = bt2.argbest()
            xb = bt2.best(ln)
            upd = bt2.counters()["updates"]
            nbytes = xb.nbytes + sum(v.nbytes for v in out2.values())
            t5 = time.perf_counter()
            bt2.close()
            pb2.close()
            t6 = time.perf_counter()
            for i, (a, b) in enumerate(zip((t0, t1, t2, t3, t4, t5), (t1, t2, t3, t4, t5, t6))):
                split[i] += 1e3 * (b - a)
            return upd, nbytes

        e2e_step(pinned_like(seeds_of(0)))  # untimed warm-up (allocator, first-use costs)
        split = [0.0] * 6
        e2e_updates = 0
        barrier()
        te = time.perf_counter()
        for s in range(args.steps):
            upd, d2h = e2e_step(step_seeds[s])
            e2e_updates += upd
        barrier()
        e2e_wall = time.perf_counter() - te
        te_t = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
        eu_t = torch.tensor([float(e2e_updates)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te_t, op=dist.ReduceOp.MAX)
            dist.all_reduce(eu_t, op=dist.ReduceOp.SUM)
        e2e = {"value": eu_t.item() / te_t.item(), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * te_t.item() / args.steps,
               "split_ms_per_step": dict(zip(("problem_create", "batch_create", "begin", "solve", "results",
                                              "destroy"), [round(v / args.steps, 2) for v in split]))}

    bt.close()
    pb.close()

    if rank == 0:
        peak, peak_kind = hbm_peak()
        # dominant kernel = k_solve (one launch per step per GPU); bytes and
        # time below are per launch on this rank
        per_launch_bytes = cnt["algorithmic_bytes"] / args.steps
        per_launch_s = solve_ms / args.steps / 1e3
        achieved = per_launch_bytes / per_launch_s / 1e9
        line = {
            "metric": METRIC, "value": updates / wall_max, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall_max / args.steps,
            "device_ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": config_of(args, wname, inst), "device_bytes": info["device_bytes"],
            "sweeps_per_s": sweeps / wall_max,
            "best": {"remaining": inc_rem, "value": inc_value if inc_rem == 0 else None, "rank": winner,
                     "ones": int(np.unpackbits(inc_x.view(np.uint8)).sum())},
            "roofline": {"bound": "hbm", "kernel": "k_solve", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "peak_kind": peak_kind,
                         "traffic": measured_traffic(args.workload, L, args.limit),
                         "bytes_per_launch": per_launch_bytes,
                         "ms_per_launch": 1e3 * per_launch_s},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
        }
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            # the whole batch once on the host cores: 10-30 s of CPU work
            n_cpu = args.cpu_lanes or min(L, 1024)
            u, _, dt, kind, threads = cpu_solve_sample(inst, args, n_cpu, 1)
            line["cpu_baseline"] = {
                "value": u / dt, "unit": UNIT, "cores": threads, "kind": kind,
                "sample": f"{n_cpu} of {L} restarts, limit={args.limit}, {threads} host threads, {dt:.1f}s",
            }
        print(json.dumps(line), flush=True)

    if world > 1:
        dist.destroy_process_group()
    return 0



# ---------------------------------------------------------------------------
# optimizer mode
# ---------------------------------------------------------------------------
def config_of(args, wname, inst):
    """The `config` object: the same keys and values in both arms."""
    nnz = int(inst["row_ptr"][-1])
    if args.mode == "optimize":
        return {"workload": wname, "mode": "optimize (optimize_functor per worker, shared population)",
                "workers_per_gpu": args.lanes, "limit": args.limit, "rounds_per_step": args.rounds,
                "exchange_every": args.exchange_every, "seed": args.seed, "m": inst["m"], "n": inst["n"],
                "nnz": nnz, "l2": "inputs larger than L2 (P is %.0f MB per GPU)" % (args.lanes * nnz * 8 / 1e6)}
    return {"workload": wname, "mode": "solve (solver_functor per restart)", "lanes_per_gpu": args.lanes,
            "limit": args.limit, "m": inst["m"], "n": inst["n"], "nnz": nnz,
            "l2": "inputs larger than L2 (P is %.0f MB per GPU)" % (args.lanes * nnz * 8 / 1e6)}


def cpu_optimize_sample(inst, args, seconds, seed):
    """The reference's own optimize_problem<> (all host threads, counting Solver
    subclass) on the workload for `seconds`: the faster of its two builds
    (-O2 parity flags without contract checks, and the
    reference's release flags -O3 + FMA).  Returns a dict."""
    from baryonyx_b200 import synth
    from oracle import pyoracle as po

    cores = os.cpu_count() or 1
    p = po.default_params()
    p.limit = args.limit
    if args.workload == "telebus":
        p.order = 3
    wanted = {"both": (True, "o3"), "o2": (True,), "o3": ("o3",)}[args.cpu_builds]
    flavours = [f for f in wanted if po.ref_available(f)] or [f for f in (True, "o3") if po.ref_available(f)]
    if not flavours:
        raise SystemExit("bench.py: oracle/_ref is not built (python -c 'import __graft_entry__ as g; g.build()')")
    runs = []
    for fl in flavours:
        pb = po.RefProblem(None, p, fast=fl, inst=inst)
        r = pb.optimize(seed, cores, seconds)
        pb.close()
        r["flavour"] = "-O3 -march=x86-64-v3, FMA, no contract checks" if fl == "o3" else \
            "-O2 -ffp-contract=off, no contract checks"
        r["rate"] = r["updates"] / r["seconds"]
        r["threads"] = cores
        del r["x"]
        runs.append(r)
    runs.sort(key=lambda r: -r["rate"])
    out = runs[0]
    out["other"] = runs[1] if len(runs) > 1 else None
    return out


def run_reference_optimize(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    inst, wname = make_instance(args.workload)
    seconds = args.cpu_seconds or 8.0
    for w in range(min(args.warmup, 1)):
        cpu_optimize_sample(inst, args, 2.0, args.seed)
    upd = swp = rst = 0
    sec = 0.0
    best = None
    r = None
    for s in range(args.steps):
        r = cpu_optimize_sample(inst, args, seconds, args.seed + s)
        upd += r["updates"]
        swp += r["sweeps"]
        rst += r.get("restarts", 0)
        sec += r["seconds"]
        best = r
    value = upd / sec
    sample = (f"optimize_problem, {r['threads']} host threads, time_limit {seconds:.0f} s per step, "
              f"{r['flavour']} (the faster build)")
    if r.get("other"):
        sample += f"; the other build: {r['other']['rate']:.3g} updates/s"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_of(args, wname, inst),
        "sweeps_per_s": swp / sec, "restarts_per_s": rst / sec,
        "time_to_feasible_s": best.get("first_feasible_s"),
        "best": {"remaining": best["remaining"], "value": best["value"] if best["remaining"] == 0 else None},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["threads"], "kind": "reference",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_engine_optimize(args):
    import torch
    import torch.distributed as dist

    from baryonyx_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU path")
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # the engine's own communicator (bxg_comm_init): rank 0 creates the NCCL id,
        # torch.distributed only carries its 128 bytes to the other ranks
        ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            ident = torch.frombuffer(bytearray(capi.Comm.unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(ident, 0)
        comm = capi.Comm(bytes(ident.cpu().numpy().tobytes()), rank, world, local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    inst, wname = make_instance(args.workload)
    cp = bench_params(args, capi)
    cp.seed = args.seed
    L = args.lanes

    def make_optimizer(instance):
        """Problem, batch, population and worker seeds as optimize_problem sets them
        up (itm-optimizer-common.hpp:793-826): every rank builds the same population
        from the master seed and keeps its own slice of the workers' seeds."""
        pb = capi.Problem(instance, device=local)
        bt = pb.batch(L, cp)
        pop = capi.Population(pb, cp, args.seed, workers=L)
        seeds = pop.generate_seeds(world * L)
        for l in range(L):
            pop.ctx(l, int(seeds[rank * L + l]))
        return pb, bt, pop

    pb, bt, pop = make_optimizer(inst)
    info = bt.info()

    def one_step():
        pop.optimize(bt, max_rounds=args.rounds, comm=comm, exchange_every=args.exchange_every)

    for s in range(args.warmup):
        one_step()

    bt.reset_counters()
    launches0 = bt.launches()
    exchanges0 = comm.exchanges() if comm else 0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        one_step()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    cnt = bt.counters()
    dev_ms = bt.device_ms()
    launches = bt.launches() - launches0
    exchanges = (comm.exchanges() - exchanges0) if comm else 0
    stats = pop.stats()
    best = pop.member(0)

    t = torch.tensor([wall, dev_ms], dtype=torch.float64, device="cuda")
    w = torch.tensor([float(cnt["updates"]), float(cnt["sweeps"]), cnt["algorithmic_bytes"], float(launches)],
                     dtype=torch.float64, device="cuda")
    fs = stats["first_solution_s"]
    f = torch.tensor([fs if fs is not None else float("inf")], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        dist.all_reduce(f, op=dist.ReduceOp.MIN)
    wall_max, dev_ms_max = t.tolist()
    updates, sweeps, alg_bytes, launches_all = w.tolist()
    first_solution = f.item()

    # ---- end to end through the C ABI with host buffers ---------------------
    e2e = None
    if not args.no_e2e:
        keys = ("row_ptr", "col", "factor", "cmin", "cmax", "obj_var", "obj_factor")
        hinst = dict(inst)
        for k in keys:
            hinst[k] = pinned_like(inst[k])
        h2d = sum(hinst[k].nbytes for k in keys)
        split = [0.0] * 4

        def e2e_step():
            """What a caller of the C ABI does: merged constraints from host memory to
            a new problem, batch and population (built on the host from the master
            seed, scored and kept on the device), the rounds, the best member and the
            population's scores back to the host."""
            t0 = time.perf_counter()
            pb2, bt2, pop2 = make_optimizer(hinst)
            t1 = time.perf_counter()
            pop2.optimize(bt2, max_rounds=args.rounds, comm=comm, exchange_every=args.exchange_every)
            t2 = time.perf_counter()
            top = pop2.member(0)
            upd = bt2.counters()["updates"]
            nbytes = top["x"].nbytes + 40 * pop2.size()
            t3 = time.perf_counter()
            pop2.close()
            bt2.close()
            pb2.close()
            t4 = time.perf_counter()
            for i, (a, b) in enumerate(zip((t0, t1, t2, t3), (t1, t2, t3, t4))):
                split[i] += 1e3 * (b - a)
            return upd, nbytes

        e2e_step()  # untimed warm-up (allocator, first-use costs)
        split = [0.0] * 4
        e2e_updates = 0
        d2h = 0
        barrier()
        te = time.perf_counter()
        for s in range(args.steps):
            upd, d2h = e2e_step()
            e2e_updates += upd
        barrier()
        e2e_wall = time.perf_counter() - te
        te_t = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
        eu_t = torch.tensor([float(e2e_updates)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te_t, op=dist.ReduceOp.MAX)
            dist.all_reduce(eu_t, op=dist.ReduceOp.SUM)
        e2e = {"value": eu_t.item() / te_t.item(), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * te_t.item() / args.steps,
               "note": "every step starts a new optimizer from x = 0 (first restarts: all rows violated)",
               "split_ms_per_step": dict(zip(("create", "rounds", "results", "destroy"),
                                             [round(v / args.steps, 2) for v in split]))}

    pop.close()
    bt.close()
    pb.close()
    if comm:
        comm.close()

    if rank == 0:
        peak, peak_kind = hbm_peak()
        n_launch = args.steps * args.rounds
        per_launch_bytes = cnt["algorithmic_bytes"] / n_launch
        per_launch_s = dev_ms / n_launch / 1e3
        achieved = per_launch_bytes / per_launch_s / 1e9
        restarts = world * L * args.steps * args.rounds
        line = {
            "metric": METRIC, "value": updates / wall_max, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall_max / args.steps,
            "device_ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_of(args, wname, inst),
            "device_bytes": info["device_bytes"],
            "sweeps_per_s": sweeps / wall_max, "restarts_per_s": restarts / wall_max,
            "algorithmic_bytes_per_s": alg_bytes / wall_max,
            "time_to_feasible_s": None if first_solution == float("inf") else first_solution,
            "exchanges_in_timed_region": int(exchanges),
            "best": {"remaining": best["remaining"], "value": best["value"] if best["remaining"] == 0 else None},
            "roofline": {"bound": "hbm", "kernel": "k_optimize", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "peak_kind": peak_kind,
                         "traffic": measured_traffic(args.workload, L, args.limit, "optimize"),
                         "bytes_per_launch": per_launch_bytes, "ms_per_launch": 1e3 * per_launch_s},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
        }
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            secs = args.cpu_seconds or 12.0
            r = cpu_optimize_sample(inst, args, secs, args.seed)
            line["cpu_baseline"] = {
                "value": r["rate"], "unit": UNIT, "cores": r["threads"], "kind": "reference",
                "sample": f"optimize_problem for {r['seconds']:.1f} s on {r['threads']} host threads, {r['flavour']}"
                          + (f"; the other build: {r['other']['rate']:.3g} updates/s" if r.get("other") else ""),
                "restarts_per_s": r.get("restarts", 0) / r["seconds"], "time_to_feasible_s": r.get("first_feasible_s"),
            }
        print(json.dumps(line), flush=True)

    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_optimize(args) if args.mode == "optimize" else run_reference(args)
    return run_engine_optimize(args) if args.mode == "optimize" else run_engine(args)


if __name__ == "__main__":
    sys.exit(main())
