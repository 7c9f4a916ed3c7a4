"""One process per GPU: read-index sharding of the hot path across ranks.

The path has no cross-read reduction (every alignment depends only on its own window and the broadcast
adaptor), so ranks take contiguous read-index ranges -- the axis .parallelize uses (R/adaptorAlign.R:126-134) --
run the C-ABI calls on their own device and the per-read results are concatenated in rank order.  No
collective touches the data path; torch.distributed (NCCL on the GPU box, gloo in the CPU tests) is used only
for the final gather of the small per-read result vectors and for barriers / max-over-ranks timing.
"""
import os

import numpy as np


def shard_bounds(n, world):
    """[lo, hi) per rank: lo_r = n*r // world, the same rule the library uses across its configured devices."""
    return [(n * r // world, n * (r + 1) // world) for r in range(world)]


def rank_world():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def gather_concat(local, dst=None):
    """Concatenate per-rank numpy arrays (leading axis = reads) in rank order on every rank (dst=None) or on `dst`."""
    import torch.distributed as dist
    rank, world = rank_world()
    if world == 1:
        return local
    parts = [None] * world
    if dst is None:
        dist.all_gather_object(parts, local)
    else:
        dist.gather_object(local, parts if rank == dst else None, dst=dst)
        if rank != dst:
            return None
    return np.concatenate(parts, axis=-1)


def adaptor_align_sharded(reads, encoding, gapopen, gapext, adaptor, sec_starts=(), sec_ends=(), align_fn=None, device=None):
    """sarlacc_adaptor_align over this rank's shard of `reads` (a ReadSet every rank holds, or can index), results
    gathered on every rank: [score, start, end, [sec_start...], [sec_width...]] for ALL reads, identical to a
    single-process call whatever the world size."""
    rank, world = rank_world()
    n = len(reads)
    lo, hi = shard_bounds(n, world)[rank]
    if align_fn is None:
        from . import native, _lib
        if device is not None:
            _lib.check(_lib.lib.sarlacc_set_devices((_lib.C.c_int * 1)(device), 1))
        align_fn = native.adaptor_align
    mine = reads[np.arange(lo, hi)]
    out = align_fn(mine, encoding, gapopen, gapext, adaptor, sec_starts, sec_ends)
    score = gather_concat(np.asarray(out[0]))
    start = gather_concat(np.asarray(out[1]))
    end = gather_concat(np.asarray(out[2]))
    nsec = len(out[3])
    sst = [gather_concat(np.asarray(out[3][s])) for s in range(nsec)]
    swd = [gather_concat(np.asarray(out[4][s])) for s in range(nsec)]
    return [score, start, end, sst, swd]
