#!/usr/bin/env python
"""Multi-GPU check of the incumbent exchange (bxg_comm_init / bxg_exchange_best), one rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_exchange.py
Every rank runs a few optimizer rounds from its own worker seeds, then the exchange: afterwards every
rank's population must hold the box-wide best member (same scores, same vector), and the winner must be
the rank whose best was best (fewest remaining constraints, then objective, then lowest rank)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from baryonyx_b200 import capi, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    ident = torch.frombuffer(bytearray(capi.Comm.unique_id()), dtype=torch.uint8).cuda()
dist.broadcast(ident, 0)
comm = capi.Comm(bytes(ident.cpu().numpy().tobytes()), rank, world, local)

inst = synth.spp(m=20, n=120, lo=3, hi=5, seed=3, cost_hi=9)
p = capi.default_params()
for k, v in dict(limit=400, pushes_limit=4, pushing_iteration_limit=6, init_population_size=20,
                 init_kappa_improve_stop=0.06, kappa_step=0.01).items():
    setattr(p, k, v)
cp = capi.clamp_params(p)
L = 4
pb = capi.Problem(inst, device=local)
bt = pb.batch(L, cp)
pop = capi.Population(pb, cp, 77, workers=L)
seeds = pop.generate_seeds(world * L)
for l in range(L):
    pop.ctx(l, int(seeds[rank * L + l]))
ok = True
for it in range(6):
    pop.optimize(bt, max_rounds=2 + rank)  # ranks drift apart
    mine = pop.member(0)
    key = torch.tensor([float(mine["remaining"]), mine["value"], float(rank)], dtype=torch.float64, device="cuda")
    keys = [torch.zeros_like(key) for _ in range(world)]
    dist.all_gather(keys, key)
    keys = [k.tolist() for k in keys]
    want = min(range(world), key=lambda r: (keys[r][0], keys[r][1], r))
    hashes = [None] * world
    dist.all_gather_object(hashes, int(mine["hash"]))
    winner = pop.exchange(bt, comm)
    ms = pop.members()
    have = any(m["hash"] == hashes[want] and m["remaining"] == int(keys[want][0]) for m in ms)
    srt = [(m["remaining"], m["value"]) for m in ms]
    good = winner == want and have and srt == sorted(srt)
    ok &= good
    if rank == 0 or not good:
        print(f"[rank {rank}] exchange {it}: winner {winner} (expected {want}), incumbent present {have}, "
              f"best now ({ms[0]['remaining']}, {ms[0]['value']})", flush=True)
flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("exchange check:", "OK" if flag.item() == 1.0 else "FAILED", f"({world} ranks, {comm.exchanges()} exchanges)")
pop.close(); bt.close(); pb.close(); comm.close()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)
