"""CPU-only, world_size 2 over gloo: the N > 1 path = contiguous read-index shards per rank + gather of the
per-read results, with no collective on the data path.  The compute function is injected (the CPU oracle here;
the CUDA library on the GPU box), so this exercises exactly the host logic bench.py and
sarlacc_b200.distributed use across ranks."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import VIGNETTE_A2, random_windows
    from oracle.oracle import Oracle, phred_encoding
    from sarlacc_b200 import ReadSet, distributed as D
    O = Oracle("port")
    enc = phred_encoding()
    rng = np.random.default_rng(99)                  # same data on every rank
    seqs, quals = random_windows(rng, 101, VIGNETTE_A2, 5, 80)
    reads = ReadSet.from_strings(seqs, quals)

    def align_fn(rs, enc, go, ge, adaptor, ss, se):
        sc, st, en, a, b = O.adaptor_align(rs.seq_strings(), rs.qual_strings(), enc, go, ge, adaptor, ss, se)
        return [sc, st, en, list(a), list(b)]

    got = D.adaptor_align_sharded(reads, enc, 5, 1, "AAGGCCTTNNNNCGACTCATGAA", [8], [12], align_fn=align_fn)
    exp = O.adaptor_align(seqs, quals, enc, 5, 1, "AAGGCCTTNNNNCGACTCATGAA", [8], [12])
    ok = (np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1]) and np.array_equal(got[2], exp[2])
          and np.array_equal(got[3][0], exp[3][0]) and np.array_equal(got[4][0], exp[4][0]))
    lo, hi = D.shard_bounds(len(reads), world)[rank]
    # max-over-ranks reduction used for timing in bench.py
    import torch
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and float(t.item()) == float(world) and (hi - lo) in (50, 51)
    with open(os.path.join(out_dir, "rank%d" % rank), "w") as fh:
        fh.write("ok" if ok else "bad")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(str(tmp_path / ("rank%d" % r))).read() == "ok"


def test_shard_bounds_tile_the_reads():
    sys.path.insert(0, ROOT)
    from sarlacc_b200.distributed import shard_bounds
    for n in (0, 1, 7, 1000003):
        for w in (1, 2, 4, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
