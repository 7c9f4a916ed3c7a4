"""configs[4] (BASELINE.json): adaptorAlign + getAdaptorThresholds on synthetic 5 kb reads sharded over the GPUs of one
box by read index.  Every rank (torchrun) or the single process takes reads [rank * share, (rank + 1) * share) in chunks:
the device generates the chunk's windows (sarlacc_chunk_load_mock), aligns both adaptors to both ends and walks back the
kept strand (sarlacc_chunk_adaptor_align: result columns stream into page-locked host tables), scrambles the windows and
scores them four more times (sarlacc_chunk_scrambled_scores: kept scores stay on the device); at the end the score vectors
of all ranks are gathered and the two thresholds selected on rank 0's device (sarlacc_compute_threshold).  Wall time
covers all of it, generation included.  A strided sample is checked against the reference's own C++ afterwards.

usage: python tools/run_c5.py [--total N] [--chunk M] [--check-stride S]      or under torchrun (one rank per GPU)"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sarlacc_b200 import native, synth, _lib  # noqa: E402

A1 = "ACGCAGATCGATCGATNNNNNNNNNNNNCGCGCGAGCTGACTNNNNGCACGACTCTGGTTTTTTTTTTTT"
A2 = "AAGGCCTTTTCCGACTCATGAA"
S1, E1 = [16, 42], [28, 46]
SEED, SCR_SEED, TOL = 5000, 1, 250


def run(total, chunk, rank, world, local, check_stride=0, error=0.01, timing=True):
    """Returns the record (dict) on every rank; thresholds / parity fields only on rank 0."""
    import torch
    import torch.distributed as dist
    share = (total + world - 1) // world
    lo, hi = rank * share, min(total, (rank + 1) * share)
    n = max(0, hi - lo)
    enc = native.phred_encoding()
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()   # noqa: E731
    nn = max(n, 1)
    out = {"reversed": pin(nn, torch.uint8), "width": pin(nn, torch.int32),
           "start1": pin(nn, torch.int32), "end1": pin(nn, torch.int32), "sec_start1": pin((2, nn), torch.int32), "sec_width1": pin((2, nn), torch.int32),
           "start2": pin(nn, torch.int32), "end2": pin(nn, torch.int32)}
    dev = torch.device("cuda", local)
    nbuf = max(nn, 8)
    real1 = torch.zeros(nbuf, dtype=torch.float64, device=dev)
    real2 = torch.zeros(nbuf, dtype=torch.float64, device=dev)
    scr1 = torch.zeros(nbuf, dtype=torch.float64, device=dev)
    scr2 = torch.zeros(nbuf, dtype=torch.float64, device=dev)
    ch = native.Chunk(min(chunk, nn), TOL, enc, device=local)
    # warm-up outside the timed region: module load, plan upload, scratch allocation
    ch.load_mock(min(chunk, nn), A1, A2, seed=SEED, first_index=lo)
    ch.adaptor_align(5, 1, A1, A2, (S1, E1), ((), ()), out={"score1": real1.data_ptr()}, out_pitch=nn)
    ch.scrambled_scores(5, 1, A1, A2, seed=SCR_SEED, first_index=lo, score1=scr1.data_ptr(), score2=scr2.data_ptr())
    ch.sync()
    gathered = [None] * 4
    if world > 1:       # receive buffers on rank 0 and one small gather, so that the timed one finds its channels set up
        gathered = [torch.zeros(world * share, dtype=torch.float64, device=dev) if rank == 0 else None for _ in range(4)]
        dist.gather(real1[:share] if n == share else torch.cat([real1[:n], real1.new_zeros(share - n)]),
                    list(gathered[0].chunk(world)) if rank == 0 else None, dst=0)     # same size as the timed ones: connections and buffers set up
        torch.cuda.synchronize()
    if rank == 0:       # loads the sort kernels and sizes the library's cached sort buffers for the job (a device allocation of this size can take 0.5 s)
        w1, w2 = (gathered[0], gathered[2]) if world > 1 else (real1, scr1)
        native.compute_threshold((w1.data_ptr(), total), (w2.data_ptr(), total), error, device=local)
    ch.set_timing(timing)
    _lib.lib.sarlacc_kernel_launches(1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for b0 in range(0, n, chunk):
        m = min(chunk, n - b0)
        ch.load_mock(m, A1, A2, seed=SEED, first_index=lo + b0)
        o = {k: (v[:, b0:] if v.ndim == 2 else v[b0:]) for k, v in out.items()}
        o["score1"] = real1.data_ptr() + 8 * b0
        o["score2"] = real2.data_ptr() + 8 * b0
        ch.adaptor_align(5, 1, A1, A2, (S1, E1), ((), ()), out=o, out_pitch=nn)
        ch.scrambled_scores(5, 1, A1, A2, seed=SCR_SEED, first_index=lo + b0, score1=scr1.data_ptr() + 8 * b0, score2=scr2.data_ptr() + 8 * b0)
    ch.sync()
    t_align = time.perf_counter() - t0
    # the one step that needs every read: gather the four score vectors, select the thresholds on rank 0.  Ranks hold
    # consecutive index ranges of `share` reads (the last ones may be short), so the first `total` entries of the gathered
    # buffers are exactly the reads of the job.
    thr = None
    if world > 1:
        for src, dst in zip((real1, real2, scr1, scr2), gathered):
            dist.gather(src[:share] if n == share else torch.cat([src[:n], src.new_zeros(share - n)]),
                        list(dst.chunk(world)) if rank == 0 else None, dst=0)
        vecs = gathered
    else:
        vecs = [real1, real2, scr1, scr2]
    t_gather = time.perf_counter() - t0 - t_align
    if rank == 0:
        thr = (native.compute_threshold((vecs[0].data_ptr(), total), (vecs[2].data_ptr(), total), error, device=local),
               native.compute_threshold((vecs[1].data_ptr(), total), (vecs[3].data_ptr(), total), error, device=local))
    t_thr = time.perf_counter() - t0 - t_align - t_gather
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_all = time.perf_counter() - t0
    launches = int(_lib.lib.sarlacc_kernel_launches(0))
    phases = ch.phase_ms() if timing else None
    tt = torch.tensor([t_all, t_align], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_all, t_align = float(tt[0]), float(tt[1])
    rec = {"reads": total, "n_gpus": world, "reads_per_gpu": share, "chunk": chunk, "seconds": t_all, "seconds_alignment_passes": t_align, "seconds_gather_rank0": t_gather, "seconds_thresholds_rank0": t_thr,
           "reads_per_s": total / t_all, "cells_per_read": 92000, "gcups_per_gpu": total * 92000 / t_all / 1e9 / world,
           "phases_ms_rank0": phases, "gpu_launches_rank0": launches, "kernels": [ch.last_kernel(0), ch.last_kernel(1)],
           "d2h_bytes_per_read": 1 + 4 + 4 * 4 + 16, "thresholds": thr}
    # parity on a strided sample: the reference's own C++ on the host mirror of the same reads
    if check_stride and rank == 0 and n > 0:
        from oracle.oracle import Oracle, phred_encoding
        O = Oracle("ref" if Oracle.available("ref") else "port")
        oenc = phred_encoding()
        pick = np.arange(0, n, check_stride)
        checked = 0
        r1 = real1.cpu().numpy()
        r2 = real2.cpu().numpy()
        for i in pick:
            f, b, w, _ = synth.mock_windows(1, A1, A2, seed=SEED, first_index=lo + int(i))
            fa, ba = ((f.seq_pool, f.seq_off), (f.qual_pool, f.qual_off)), ((b.seq_pool, b.seq_off), (b.qual_pool, b.qual_off))
            a = O.adaptor_align(*fa, oenc, 5, 1, A1, S1, E1)
            bb = O.adaptor_align(*ba, oenc, 5, 1, A2)
            c = O.adaptor_align(*ba, oenc, 5, 1, A1, S1, E1)
            d = O.adaptor_align(*fa, oenc, 5, 1, A2)
            rev = (max(a[0][0], 0) + max(bb[0][0], 0)) < (max(c[0][0], 0) + max(d[0][0], 0))
            x, y = (c, d) if rev else (a, bb)
            assert bool(out["reversed"][i]) == rev and out["width"][i] == w[0]
            assert r1[i] == x[0][0] and r2[i] == y[0][0]
            assert out["start1"][i] == x[1][0] and out["end1"][i] == x[2][0]
            assert out["start2"][i] == w[0] - y[1][0] + 1 and out["end2"][i] == w[0] - y[2][0] + 1
            for s in range(2):
                assert out["sec_start1"][s, i] == x[3][s][0] and out["sec_width1"][s, i] == x[4][s][0]
            checked += 1
        rec["parity"] = {"checked_reads": checked, "stride": check_stride, "oracle": O.kind, "identical": True}
    ch.close()
    return rec


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--total", type=int, default=6250000)
    ap.add_argument("--chunk", type=int, default=1136640)
    ap.add_argument("--check-stride", type=int, default=0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rec = run(args.total, args.chunk, rank, world, local, args.check_stride)
    if rank == 0:
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        peak = sms * 64 * 1.965e9 / 10 / 1e9
        rec["roofline_frac"] = rec["gcups_per_gpu"] / peak
        print(json.dumps(rec))
    if world > 1:
        torch.distributed.destroy_process_group()
