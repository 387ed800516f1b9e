"""Incumbent exchange between the ranks of one box.

The path shards over restarts: every rank (one process per GPU) owns its lanes
and a replica of the problem structure; there is no data-path collective.  The
only exchange is the one the reference's workers do through the shared
best_solution_recorder (lib/src/itm-optimizer-common.hpp:461-594): who holds
the best solution, and its bits.

    key   = (remaining constraints, objective in the order of
             is_better_solution<Mode>, rank)            -> all_reduce(MIN)
    x     = the winner's bit_array words                -> broadcast(src=winner)

`torch.distributed` is the plumbing: NCCL over NVLink on the GPU box (device
tensors), gloo in the CPU tests (host tensors).  Ties are broken by the lower
rank, so every rank ends with the same incumbent, bit for bit.
"""
from __future__ import annotations

import numpy as np


def shard_lanes(total: int, world: int, rank: int):
    """Lanes [first, first + count) of `total` restarts owned by `rank`: an
    even split, the remainder going to the low ranks."""
    base, extra = divmod(int(total), int(world))
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


_M = 2147483647  # modulus of minstd_rand0
_U64 = (1 << 64) - 1


def _uniform_u64(state: int, urange: int):
    """libstdc++ 13 uniform_int_distribution<unsigned long>(0, urange) on
    minstd_rand0 (bits/uniform_int_dist.h): scale-and-reject when the engine's
    range (2^31 - 3) covers urange, recursive upscaling in 64-bit wrap-around
    arithmetic otherwise.  Returns (value, new engine state)."""
    urng_range = _M - 2
    if urng_range > urange:
        uerange = urange + 1
        scaling = urng_range // uerange
        past = uerange * scaling
        while True:
            state = (16807 * state) % _M
            r = state - 1
            if r < past:
                return r // scaling, state
    if urng_range < urange:
        uerng = urng_range + 1
        while True:
            hi, state = _uniform_u64(state, urange // uerng)
            tmp = (uerng * hi) & _U64
            state = (16807 * state) % _M
            ret = (tmp + (state - 1)) & _U64
            if not (ret > urange or ret < tmp):
                return ret, state
    state = (16807 * state) % _M
    return state - 1, state


def lane_seeds(master_seed: int, total: int, first: int = 0, count: int | None = None) -> np.ndarray:
    """Engine seeds of `total` restarts as generate_seed (lib/src/itm-common.hpp
    :1457-1484) hands them to the optimizer's workers: 64-bit uniform draws from
    the master engine random_engine(seed) (itm-optimizer-common.hpp:793,817),
    re-drawn while two are equal.  A worker seeds minstd_rand0 with its 64-bit
    value, i.e. with value mod (2^31 - 1): that residue is what is returned
    (bxg_batch_begin takes 32-bit seeds).  Every rank replays the master stream
    and keeps lanes [first, first + count), so the seeds of a restart do not
    depend on how many GPUs share the work."""
    if count is None:
        count = total - first
    state = int(master_seed) % _M or 1
    ret = [0] * total
    if total:
        ret[0], state = _uniform_u64(state, _U64)
    for i in range(1, total):
        ret[i], state = _uniform_u64(state, _U64)
        j = i - 1
        while j > 0:
            if ret[j] == ret[i]:
                ret[j], state = _uniform_u64(state, _U64)
                j = i
            j -= 1
    return np.array([v % _M for v in ret[first:first + count]], dtype=np.uint32)


def better_key(remaining: int, value: float, maximize: bool):
    """Total order of raw_result (itm-optimizer-common.hpp:427-456):
    fewer remaining constraints first, then the better objective."""
    if remaining != 0:
        return float(remaining), 0.0
    return 0.0, (-value if maximize else value)


def exchange_best(remaining: int, value: float, x_words: np.ndarray, maximize: bool = False,
                  group=None, device=None):
    """All ranks call this with their own best (remaining constraints, value,
    bit_array words).  Returns (winner_rank, remaining, value, x_words) of the
    box-wide incumbent, identical on every rank."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0, int(remaining), float(value), np.array(x_words, copy=True)

    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")

    k0, k1 = better_key(remaining, value, maximize)
    # three lexicographic MIN reductions: remaining, objective, rank
    a = torch.tensor([k0], dtype=torch.float64, device=dev)
    dist.all_reduce(a, op=dist.ReduceOp.MIN, group=group)
    mine = k0 == a.item()
    b = torch.tensor([k1 if mine else float("inf")], dtype=torch.float64, device=dev)
    dist.all_reduce(b, op=dist.ReduceOp.MIN, group=group)
    mine = mine and (k1 == b.item() or (k1 != k1 and b.item() != b.item()))
    c = torch.tensor([rank if mine else world], dtype=torch.int64, device=dev)
    dist.all_reduce(c, op=dist.ReduceOp.MIN, group=group)
    winner = int(c.item())
    if winner >= world:  # every objective was NaN: rank 0 keeps its own
        winner = 0

    xw = np.ascontiguousarray(x_words, np.uint64)
    t = torch.from_numpy(xw.view(np.int64).copy()).to(dev)
    sv = torch.tensor([float(remaining), float(value)], dtype=torch.float64, device=dev)
    src = dist.get_global_rank(group, winner) if group is not None else winner
    dist.broadcast(t, src=src, group=group)
    dist.broadcast(sv, src=src, group=group)
    out = t.cpu().numpy().view(np.uint64)
    return winner, int(sv[0].item()), float(sv[1].item()), out
