"""Multi-GPU host logic on CPU: lane sharding, the seeds of
generate_seed, and the incumbent exchange over a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest

from baryonyx_b200 import exchange
from oracle import pyoracle as po

M = 2147483647


def test_shard_lanes_covers_everything():
    for total in (0, 1, 7, 1024, 1025):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                f, c = exchange.shard_lanes(total, world, r)
                seen += list(range(f, f + c))
            assert seen == list(range(total))


def test_uniform_u64_matches_oracle():
    # the oracle's uniform_int restatement is pinned to libstdc++ in test_oracle_vs_ref.py
    L = po.oracle()
    import ctypes as C

    for seed in (1, 42, 123456789, M - 1):
        for urange in (1, 5, 99, M - 3, M - 2, M - 1, 2**33, 2**62 + 12345, 2**64 - 1):
            st = C.c_uint32(seed)
            want = [L.ox_uniform_u64(0, urange, C.byref(st)) for _ in range(6)]
            state = seed
            got = []
            for _ in range(6):
                v, state = exchange._uniform_u64(state, urange)
                got.append(v)
            assert got == want, (seed, urange)
            assert state == st.value


@pytest.mark.ref
def test_lane_seeds_match_generate_seed():
    R = po.ref()
    for master in (1, 42, 987654321, 2**31, 2**40 + 17):
        for count in (1, 2, 16, 300):
            out = np.zeros(count, np.uint64)
            R.bxref_generate_seed(master, count, out)
            want = (out % np.uint64(M)).astype(np.uint32)
            got = exchange.lane_seeds(master, count)
            assert np.array_equal(got, want), (master, count)
            # a rank's slice does not depend on the number of ranks
            f, c = exchange.shard_lanes(count, 3, 1)
            assert np.array_equal(exchange.lane_seeds(master, count, f, c), want[f:f + c])


def test_lane_seeds_golden():
    # generate_seed of the compiled reference, master seed 42 (tests/golden/make_golden.py)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "generate_seed_42.npy"))
    assert np.array_equal(exchange.lane_seeds(42, len(g)), g)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cases, ret):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = []
    for case in cases:
        rem, val, seed, maximize = case[rank]
        x = np.random.default_rng(seed).integers(0, 2**63, 9, dtype=np.uint64)
        w, r, v, xw = exchange.exchange_best(rem, val, x, maximize=maximize)
        out.append((w, r, v, xw.tolist()))
    ret[rank] = out
    dist.destroy_process_group()


def test_exchange_best_gloo_world2():
    import torch.multiprocessing as mp

    # per case: (remaining, value, x seed, maximize) of rank 0 and rank 1
    cases = [
        [(0, 10.0, 1, False), (0, 7.5, 2, False)],      # both feasible: lower value wins
        [(0, 10.0, 3, True), (0, 7.5, 4, True)],        # maximize: higher value wins
        [(3, 1.0, 5, False), (0, 99.0, 6, False)],      # feasible beats infeasible
        [(4, 0.0, 7, False), (2, 0.0, 8, False)],       # fewer remaining constraints
        [(0, 5.0, 9, False), (0, 5.0, 10, False)],      # tie: lower rank
    ]
    winners = [1, 0, 1, 1, 0]
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, cases, ret), nprocs=world, join=True)
    assert ret[0] == ret[1]  # every rank holds the same incumbent
    for i, case in enumerate(cases):
        w, r, v, xw = ret[0][i]
        assert w == winners[i], i
        rem, val, seed, _ = case[w]
        assert (r, v) == (rem, val)
        want = np.random.default_rng(seed).integers(0, 2**63, 9, dtype=np.uint64).tolist()
        assert xw == want
