"""configs[4] (BASELINE.json): adaptorAlign + getAdaptorThresholds on 50 M synthetic 5 kb reads sharded over the GPUs of
one box by read index.  Every rank (torchrun) or the single process takes `--share` reads (default 50 M / 8 = 6.25 M, the
per-GPU share of the 8-GPU job) in batches of `--batch`, keeps the packed windows resident, runs the four alignments with
traceback (adaptorAlign) and the four score-only alignments on device-scrambled windows (getAdaptorThresholds), and the
thresholds are selected from all ranks' scores on rank 0.  Synthetic read generation is outside the timed phases (it
stands in for the FASTQ file).  A strided sample of every batch is checked against the reference's own C++.

usage: python tools/run_c5.py [--share N] [--batch M]      or under torchrun (one rank per GPU)"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sarlacc_b200 import api, native, synth  # noqa: E402

A1 = "ACGCAGATCGATCGATNNNNNNNNNNNNCGCGCGAGCTGACTNNNNGCACGACTCTGGTTTTTTTTTTTT"
A2 = "AAGGCCTTTTCCGACTCATGAA"
ap = argparse.ArgumentParser()
ap.add_argument("--share", type=int, default=6250000)
ap.add_argument("--batch", type=int, default=1250000)
ap.add_argument("--check-stride", type=int, default=5003)
args = ap.parse_args()
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
import torch  # noqa: E402
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
enc = native.phred_encoding()
s1, e1 = [16, 42], [28, 46]
oracle = None
try:
    from oracle.oracle import Oracle
    oracle = Oracle("ref" if Oracle.available("ref") else "port")
except Exception:
    pass

t_align = t_thr = t_gen = 0.0
real1, real2, scr1, scr2 = [], [], [], []
checked = 0
stream = torch.cuda.Stream(device=local)
bufs = {}
# one untimed warm-up batch first (module load, page-locked result buffers), then the share in batches
for bi, b0 in enumerate([-1] + list(range(0, args.share, args.batch))):
    warm = b0 < 0
    b0 = max(b0, 0)
    m = min(args.batch, args.share - b0)
    first = rank * args.share + b0
    t0 = time.perf_counter()
    front, back, widths, _ = synth.mock_windows(m, A1, A2, seed=5000, first_index=first)
    if not warm:
        t_gen += time.perf_counter() - t0
    rf, rb = native.Resident(front, enc, device=local), native.Resident(back, enc, device=local)
    torch.cuda.synchronize()
    # ---- adaptorAlign: .align_AA_internal's four alignments with traceback + strand resolution
    t0 = time.perf_counter()
    res = {}
    for key, r, a, sec in (("a", rf, A1, (s1, e1)), ("b", rb, A2, ((), ())), ("c", rb, A1, (s1, e1)), ("d", rf, A2, ((), ()))):
        r.align(r.MODE_TRACE_LOCAL, 5, 1, a, *sec, stream=stream.cuda_stream)
        res[key] = bufs[key] = r.fetch(stream=stream.cuda_stream, pinned=True, out=bufs.get(key))
    torch.cuda.synchronize()
    rev = api._resolve_strand(res["a"][0], res["b"][0], res["c"][0], res["d"][0])["reversed"]
    if not warm:
        t_align += time.perf_counter() - t0
        real1.append(np.where(rev, res["c"][0], res["a"][0]))
        real2.append(np.where(rev, res["d"][0], res["b"][0]))
    # ---- getAdaptorThresholds: scramble on the device (keyed by global read index), four score-only alignments
    t0 = time.perf_counter()
    idx = np.arange(first, first + m, dtype=np.uint64)
    sf, sb = rf.scrambled(0, read_index=idx, stream_id=0), rb.scrambled(0, read_index=idx, stream_id=1)
    sc = {}
    for key, r, a in (("S", sf, A1), ("E", sb, A2), ("RS", sb, A1), ("RE", sf, A2)):
        r.align(r.MODE_SCORE_LOCAL, 5, 1, a, stream=stream.cuda_stream)
        sc[key] = bufs[key] = r.fetch(stream=stream.cuda_stream, pinned=True, out=bufs.get(key))
    torch.cuda.synchronize()
    srev = api._resolve_strand(sc["S"], sc["E"], sc["RS"], sc["RE"])["reversed"]
    if not warm:
        t_thr += time.perf_counter() - t0
        scr1.append(np.where(srev, sc["RS"], sc["S"]))
        scr2.append(np.where(srev, sc["RE"], sc["E"]))
    # ---- parity on a strided sample: the reference's own C++ on the same windows
    if oracle is not None and not warm:
        pick = np.arange(0, m, args.check_stride)
        sub = front[pick]
        exp = oracle.adaptor_align((sub.seq_pool, sub.seq_off), (sub.qual_pool, sub.qual_off), enc, 5, 1, A1, s1, e1, nthreads=os.cpu_count() or 1)
        got = res["a"]
        assert np.array_equal(got[0][pick], exp[0]) and np.array_equal(got[1][pick], exp[1]) and np.array_equal(got[2][pick], exp[2])
        for k in range(2):
            assert np.array_equal(got[3][k][pick], exp[3][k]) and np.array_equal(got[4][k][pick], exp[4][k])
        checked += len(pick)
    for r in (rf, rb, sf, sb):
        r.close()

real1, real2, scr1, scr2 = (np.concatenate(x) for x in (real1, real2, scr1, scr2))
t0 = time.perf_counter()
if world > 1:
    def gather(x):
        t = torch.from_numpy(x).cuda()
        out = torch.empty(world * len(x), dtype=t.dtype, device="cuda") if rank == 0 else None
        dist.gather(t, list(out.chunk(world)) if rank == 0 else None, dst=0)
        return out.cpu().numpy() if rank == 0 else None
    real1, real2, scr1, scr2 = gather(real1), gather(real2), gather(scr1), gather(scr2)
    stats = torch.tensor([t_align, t_thr, t_gen], dtype=torch.float64, device="cuda")
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    t_align, t_thr, t_gen = (float(x) for x in stats.cpu())
if rank == 0:
    thr1 = api._compute_threshold(real1, scr1, 0.01)
    thr2 = api._compute_threshold(real2, scr2, 0.01)
    t_sel = time.perf_counter() - t0
    n = world * args.share
    cells = n * 46000
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    peak = sms * 64 * 1.965e9 / 10 / 1e9
    print("configs[4] share: %d reads on %d GPU(s) (%d per GPU, batches of %d)" % (n, world, args.share, args.batch))
    print("  adaptorAlign (4 alignments + traceback, resident windows, results fetched): %.3f s = %.2f M reads/s, %.0f GCUPS per GPU (%.1f %% of %.0f)"
          % (t_align, n / t_align / 1e6, cells / t_align / 1e9 / world, 100 * cells / t_align / 1e9 / world / peak, peak))
    print("  getAdaptorThresholds core (device scramble + 4 score-only alignments, scores fetched): %.3f s = %.2f M reads/s, %.0f GCUPS per GPU"
          % (t_thr, n / t_thr / 1e6, cells / t_thr / 1e9 / world))
    both = t_align + t_thr
    print("  both: %.3f s, %.0f GCUPS per GPU = %.1f %% of the FP64 roofline; threshold selection on the host %.2f s -> adaptor1 %.4f, adaptor2 %.4f"
          % (both, 2 * cells / both / 1e9 / world, 100 * 2 * cells / both / 1e9 / world / peak, t_sel, thr1, thr2))
    print("  parity: %d strided alignments per rank identical to the %s oracle; synthetic read generation (not timed above) %.1f s"
          % (checked, oracle.kind if oracle else "no", t_gen))
if world > 1:
    dist.destroy_process_group()
