"""Host-to-device copy bandwidth with all ranks copying at once (torchrun): what the platform gives the end-to-end path,
which uploads 1 044 bytes per read.  Every rank copies a 1 GiB page-locked buffer to its device `reps` times; rates are
reported per rank, alone (ranks take turns) and all together.
usage: python -m torch.distributed.run --nproc-per-node N tools/h2d_probe.py"""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
host.fill_(1)
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()


def rate(reps=8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    return reps * n / (time.perf_counter() - t0) / 1e9


alone = torch.zeros(world, dtype=torch.float64, device="cuda")
for r in range(world):
    if world > 1:
        dist.barrier()
    if r == rank:
        alone[r] = rate()
if world > 1:
    dist.barrier()
together = torch.zeros(world, dtype=torch.float64, device="cuda")
together[rank] = rate()
if world > 1:
    dist.all_reduce(alone)
    dist.all_reduce(together)
if rank == 0:
    print("H2D GB/s per rank, one rank at a time :", [round(x, 1) for x in alone.tolist()])
    print("H2D GB/s per rank, all ranks together :", [round(x, 1) for x in together.tolist()], "sum %.1f" % float(together.sum()))
if world > 1:
    dist.destroy_process_group()
