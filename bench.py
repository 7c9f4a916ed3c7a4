#!/usr/bin/env python
"""bench.py -- adaptorAlign throughput on B200 (BASELINE.json configs[1]) vs the reference's CPU path, with the other
configs as sub-records of the same JSON line.

One "step" = adaptorAlign's hot path over one batch of synthetic mockReads-style reads: the four local-global alignments
.align_AA_internal performs per read (adaptor1 x front, adaptor2 x back, adaptor1 x back, adaptor2 x front;
R/adaptorAlign.R:186-189), .resolve_strand, and the traceback of the strand that is kept -- tolerance 250, vignette
adaptors (70 bp with two N-runs, 22 bp), go=5, ge=1  ->  46 000 DP cells per read.

  value   : reads/s with the packed windows resident in HBM (sarlacc_chunk_adaptor_align; CUDA events on the chunk's stream)
  e2e     : the same work through the host-buffer C-ABI call sarlacc_adaptor_align_windows (pinned host windows -> H2D ->
            device packer -> kernels -> D2H), i.e. what the R wrapper would see
  roofline: the dominant kernel (forward pass of adaptor1, with traceback records) against the FP64 ALU roofline
            SMs * 64 lanes * f / 10 FP64 ops per cell, plus its HBM side
  cpu_baseline / --impl reference: the reference's own reference_align.cpp (oracle/_ref, compiled verbatim) on all host
            cores, on a bounded sample of the same reads
  c3, c4, c5: configs[2] (getAdaptorThresholds on the same reads), configs[3] (barcodeAlign, 96 x 24-bp barcodes) and
            configs[4] (adaptorAlign + getAdaptorThresholds on --c5-reads reads split over the ranks by read index)

Synthetic reads come from the device generator (sarlacc_chunk_load_mock); sarlacc_b200/synth.py: mock_windows is its host
mirror (same reads bit for bit) and feeds the CPU legs.

Launch: `python bench.py --gpus N --steps K --warmup W`, or under torchrun for N > 1 (one rank per GPU, reads sharded by
index, no collective on the alignment path; weak scaling for the headline: every rank aligns --reads reads).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

A1 = "ACGCAGATCGATCGATNNNNNNNNNNNNCGCGCGAGCTGACTNNNNGCACGACTCTGGTTTTTTTTTTTT"   # vignettes/correction.Rmd:41
A2 = "AAGGCCTTTTCCGACTCATGAA"                                                   # vignettes/correction.Rmd:42
GO, GE, TOL = 5.0, 1.0, 250
FP64_OPS_PER_CELL = 10      # 5 add/sub + 5 compares, SURVEY.md 8(d)
FP64_LANES_PER_SM = 64
SEED = 2000


def setup_subseqs(adaptor):
    import re
    st, en = [], []
    for m in re.finditer("[^ACTG]+", adaptor):
        st.append(m.start())      # 0-based start, as passed to the .Call (R/adaptorAlign.R:158)
        en.append(m.end())        # 1-based end
    return st, en


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(front, back, nreads, cores, repeats=1):
    """Reads/s of the reference's own C++ (oracle/_ref) on `cores` threads over the first nreads reads: the
    four adaptor_align calls of .align_AA_internal."""
    from oracle.oracle import Oracle, phred_encoding
    kind = "reference" if Oracle.available("ref") else "port"
    O = Oracle("ref" if kind == "reference" else "port")
    enc = phred_encoding()
    W = TOL

    def sub(rs):
        return (rs.seq_pool[:nreads * W], rs.seq_off[:nreads + 1]), (rs.qual_pool[:nreads * W], rs.qual_off[:nreads + 1])

    (fs, fq), (bs, bq) = sub(front), sub(back)
    s1, e1 = setup_subseqs(A1)
    s2, e2 = setup_subseqs(A2)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        O.adaptor_align(fs, fq, enc, GO, GE, A1, s1, e1, nthreads=cores)
        O.adaptor_align(bs, bq, enc, GO, GE, A2, s2, e2, nthreads=cores)
        O.adaptor_align(bs, bq, enc, GO, GE, A1, s1, e1, nthreads=cores)
        O.adaptor_align(fs, fq, enc, GO, GE, A2, s2, e2, nthreads=cores)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return kind, nreads / best, best


def cpu_score_only_rate(front, back, nreads, cores):
    """The four score-only calls of .get_alignment_scores (R/tuneAlignment.R:99-112) on the reference's C++: the CPU side of
    getAdaptorThresholds per read, without the R scramble loop."""
    from oracle.oracle import Oracle, phred_encoding
    kind = "reference" if Oracle.available("ref") else "port"
    O = Oracle("ref" if kind == "reference" else "port")
    enc = phred_encoding()
    W = TOL
    fa = (front.seq_pool[:nreads * W], front.seq_off[:nreads + 1]), (front.qual_pool[:nreads * W], front.qual_off[:nreads + 1])
    ba = (back.seq_pool[:nreads * W], back.seq_off[:nreads + 1]), (back.qual_pool[:nreads * W], back.qual_off[:nreads + 1])
    t0 = time.perf_counter()
    for (s, q), a in ((fa, A1), (ba, A2), (ba, A1), (fa, A2)):
        O.align_score_only(s, q, enc, GO, GE, a, nthreads=cores)
    return kind, nreads / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=1_000_000, help="reads per GPU per step (configs[1]: 1M x 5 kb)")
    ap.add_argument("--e2e-reads", type=int, default=0, help="reads per step for the host-buffer leg (0 = same as --reads)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target duration of the cpu_baseline sample")
    ap.add_argument("--c5-reads", type=int, default=50_000_000, help="configs[4]: reads of the whole job, split over the ranks")
    ap.add_argument("--c4-sequences", type=int, default=1_000_000, help="configs[3]: sequences per GPU")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the c3 / c4 / c5 sub-records")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    cells_per_read = 2 * TOL * (len(A1) + len(A2))
    workload = "configs[1]: %d synthetic 5 kb mockReads-style reads per GPU, adaptorAlign both ends " \
               "(4 local-global alignments + traceback per read), tolerance %d, vignette adaptors" % (args.reads, TOL)

    from sarlacc_b200 import synth

    # ------------------------------------------------------------------ reference arm (CPU) -------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        # bounded sample per step: ~3 s of all-core CPU work (0.02 GCUPS/core measured at survey time)
        per_step = int(max(256, min(args.reads, 3.0 * 0.02e9 * cores / cells_per_read)))
        front, back, _, _ = synth.mock_windows(per_step, A1, A2, tolerance=TOL, seed=SEED)
        times = []
        kind = "port"
        for it in range(args.warmup + args.steps):
            kind, rate, dt = cpu_reference_rate(front, back, per_step, cores)
            if it >= args.warmup:
                times.append(dt)
        ms = 1000.0 * sum(times) / len(times)
        value = per_step / (ms / 1000.0)
        line = {
            "impl": "reference", "metric": "adaptorAlign reads/s", "value": value, "unit": "reads/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "gcups": value * cells_per_read / 1e9,
            "config": {"workload": workload, "sample": "%d reads per step" % per_step},
            "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": "reference" if kind == "reference" else "port",
                             "sample": "first %d reads of the workload per step, all 4 alignments + traceback" % per_step},
            "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm -------------------------
    import torch
    import torch.distributed as dist
    from sarlacc_b200 import native, _lib, ReadSet

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_ranks(values):
        if world == 1:
            return [values]
        t = torch.tensor(values, dtype=torch.float64, device="cuda")
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [o.cpu().tolist() for o in out]

    n = args.reads
    dev = local_rank
    enc = native.phred_encoding()
    s1, e1 = setup_subseqs(A1)
    s2, e2 = setup_subseqs(A2)
    # every rank generates its own shard of one big read set on its device (reads keyed by global index)
    ch = native.Chunk(n, TOL, enc, device=dev)
    ch.load_mock(n, A1, A2, seed=SEED, first_index=rank * n)
    ch.sync()
    tdev = torch.device("cuda", dev)
    dout = {"reversed": torch.empty(n, dtype=torch.uint8, device=tdev), "width": torch.empty(n, dtype=torch.int32, device=tdev)}
    for k, ns in ((1, len(s1)), (2, len(s2))):
        dout["score%d" % k] = torch.empty(n, dtype=torch.float64, device=tdev)
        dout["start%d" % k] = torch.empty(n, dtype=torch.int32, device=tdev)
        dout["end%d" % k] = torch.empty(n, dtype=torch.int32, device=tdev)
        dout["sec_start%d" % k] = torch.empty((max(ns, 1), n), dtype=torch.int32, device=tdev)
        dout["sec_width%d" % k] = torch.empty((max(ns, 1), n), dtype=torch.int32, device=tdev)
    dptr = {k: v.data_ptr() for k, v in dout.items()}
    est = torch.cuda.ExternalStream(ch.stream(), device=tdev)      # the chunk's compute stream: events are recorded on it

    def step():
        ch.adaptor_align(GO, GE, A1, A2, (s1, e1), (s2, e2), out=dptr, out_pitch=n)

    for _ in range(args.warmup):
        step()
    ch.sync()
    _lib.lib.sarlacc_kernel_launches(1)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(est)
    for _ in range(args.steps):
        step()
    ch.join()                 # the compute stream waits for the tracebacks and result copies of the last step
    ev1.record(est)
    ch.sync()
    torch.cuda.synchronize()
    barrier()
    launches = int(_lib.lib.sarlacc_kernel_launches(0))
    ms_step = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
    clocks = sampler.stop() if sampler else None
    kernel_names = [ch.last_kernel(0), ch.last_kernel(1)]
    value = world * n / (ms_step / 1000.0)
    gcups = value * cells_per_read / 1e9

    # host copies of the same reads (decoded from the device rows: the host mirror would take minutes for 1 M reads)
    rf, lf, widths, _ = ch.rows(0)
    front = synth.unpack_rows(rf, lf)
    rb, lb, _, _ = ch.rows(1)
    back = synth.unpack_rows(rb, lb)
    del rf, rb

    # dominant kernel, timed alone with events inside the library (same stream), outside the timed region so that
    # the event synchronisation does not serialise it
    res_f = native.Resident(front, enc, device=dev)
    os.environ["SARLACC_NO_OVERLAP"] = "1"     # one launch over all windows, no traceback beside it
    fwd_ms = []
    T = native.Resident.MODE_TRACE_LOCAL
    res_f.set_timing(True)
    for _ in range(max(3, args.steps) + 1):
        res_f.align(T, GO, GE, A1, s1, e1)
        fwd_ms.append(res_f.forward_ms())
    del os.environ["SARLACC_NO_OVERLAP"]
    dom_kernel = res_f.last_kernel()
    fwd = statistics.median(fwd_ms[1:])
    cells_a1 = res_f.cells(len(A1))
    rows_bytes = res_f.nbytes()
    res_f.close()

    # ------------------------------------------------------------------ e2e: host buffers through the C ABI
    e2e = None
    if not args.no_e2e:
        ne = args.e2e_reads or n
        sub_f = front if ne == n else front[np.arange(ne)]
        sub_b = back if ne == n else back[np.arange(ne)]
        _lib.lib.sarlacc_set_devices((_lib.C.c_int * 1)(dev), 1)
        pageable_f, pageable_b = sub_f, sub_b

        def pinned(rs):
            # the step's inputs live in pinned host memory (the bench contract); the library then DMAs each chunk's byte
            # range straight out of these pools and packs on the device
            pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()   # noqa: E731
            return ReadSet(pin(rs.seq_pool), rs.seq_off, pin(rs.qual_pool), rs.qual_off, rs.names)

        inputs_pinned = True
        try:
            sub_f, sub_b = pinned(sub_f), pinned(sub_b)
        except RuntimeError:          # page-locking refused (memlock limit): the library stages the bytes itself
            inputs_pinned = False
        sub_w = widths[:ne].astype(np.int32)
        keep = {}

        def e2e_step(f=None, b=None):
            # .align_AA_internal + adaptor2 flip in one C-ABI call: windows uploaded and packed once, four forward passes,
            # strand resolution, traceback of the kept strand, result columns copied back
            return native.adaptor_align_windows(f or sub_f, b or sub_b, enc, GO, GE, A1, A2, (s1, e1), (s2, e2), read_width=sub_w, reuse=keep)

        def e2e_step_unfused():
            a = native.adaptor_align(sub_f, enc, GO, GE, A1, s1, e1)
            b = native.adaptor_align(sub_b, enc, GO, GE, A2, s2, e2)
            c = native.adaptor_align(sub_b, enc, GO, GE, A1, s1, e1)
            d = native.adaptor_align(sub_f, enc, GO, GE, A2, s2, e2)
            return (np.maximum(a[0], 0) + np.maximum(b[0], 0)) < (np.maximum(c[0], 0) + np.maximum(d[0], 0))

        e2e_step()
        e2e_step()
        barrier()
        torch.cuda.synchronize()
        reps = max(2, args.steps)
        phases = []
        t0 = time.perf_counter()
        for _ in range(reps):
            e2e_step()
            phases.append(native.last_pair_timing())
        torch.cuda.synchronize()
        dt = max_over_ranks((time.perf_counter() - t0) / reps)
        ph = {k: statistics.mean(p[k] for p in phases) for k in phases[0]}
        ph_ranks = gather_ranks([ph["stage"], ph["enqueue"], ph["wait_copy_out"], ph["total"], ph["upload_sum"], ph["kernels_copyback_sum"], ph["upload_bytes"]])
        # plain bytes: bases + qualities of both windows, two 8-byte offsets and a 4-byte length per window, read widths.
        # What the library actually copied (it counts the window data; lengths and widths added here) is the same unless
        # SARLACC_PACK_SEQ=1 makes it send the bases as 4-bit codes (include/sarlacc_b200.h; off by default).
        h2d_plain = int(2 * (sub_f.seq_off[-1] + sub_b.seq_off[-1]) + 2 * ne * (16 + 4) + ne * 4)
        h2d_ranks = [int(r[6]) + 2 * ne * 4 + ne * 4 for r in ph_ranks]
        h2d = max(h2d_ranks)
        d2h = ne * (1 + (8 + 4 + 4 + 8 * len(s1)) + (8 + 4 + 4 + 8 * len(s2)))
        e2e_step_unfused()         # warm: the single-adaptor entry has its own scratch
        barrier()
        t0 = time.perf_counter()
        e2e_step_unfused()
        torch.cuda.synchronize()
        dt_unfused = max_over_ranks(time.perf_counter() - t0)
        # the same call with ordinary (pageable) host buffers: the library gathers the bytes into its own pinned staging
        e2e_step(pageable_f, pageable_b)
        barrier()
        t0 = time.perf_counter()
        e2e_step(pageable_f, pageable_b)
        torch.cuda.synchronize()
        dt_pageable = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * ne / dt, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "h2d_bytes_per_step_plain": h2d_plain, "h2d_bytes_per_step_per_rank": h2d_ranks,
               "reads_per_step": ne, "ms_per_step": dt * 1000.0, "steps": reps,
               "path": "sarlacc_adaptor_align_windows: pinned host CSR buffers -> H2D of the raw bytes -> device packer -> 4 forward passes + "
                       "strand resolution + traceback of the kept strand on device -> D2H of the result columns (output arrays reused between calls)",
               "inputs_pinned": inputs_pinned, "pageable_inputs_reads_per_s": world * ne / dt_pageable,
               "host_threads_per_rank": max(1, cores // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))),
               "phases_ms_per_rank": [dict(zip(("host_stage", "host_enqueue", "host_wait_copy_out", "host_total", "device_upload_sum", "device_kernels_copyback_sum"), r)) for r in ph_ranks],
               "upload_gbs_per_rank": [hb / 1e6 / r[4] if r[4] > 0 else None for hb, r in zip(h2d_ranks, ph_ranks)],
               "unfused_reads_per_s": world * ne / dt_unfused,
               "unfused_path": "4 x sarlacc_adaptor_align (the reference's four .Calls) + .resolve_strand on the host"}

    # ------------------------------------------------------------------ cpu baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = int(max(256, min(n, args.cpu_seconds * 0.02e9 * cores / cells_per_read)))
        kind, rate, dt = cpu_reference_rate(front, back, sample, cores)
        cpu = {"value": rate, "unit": "reads/s", "cores": cores, "kind": kind, "seconds": dt,
               "gcups": rate * cells_per_read / 1e9,
               "sample": "first %d reads of the workload, all 4 alignments + traceback, %d threads" % (sample, cores)}

    props = torch.cuda.get_device_properties(dev)
    sms = props.multi_processor_count
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    max_mhz = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    peak_gcups = sms * FP64_LANES_PER_SM * max_mhz * 1e6 / FP64_OPS_PER_CELL / 1e9

    # ------------------------------------------------------------------ configs[2], configs[3], configs[4]
    extra = {}
    if not args.no_extra:
        extra = extra_records(args, ch, front, back, dptr, dout, enc, rank, world, local_rank, cores, peak_gcups, barrier, max_over_ranks)
    ch.close()

    if rank == 0:
        cur_mhz = (clocks or {}).get("sm_mhz") or max_mhz
        peak_gcups_at_clock = sms * FP64_LANES_PER_SM * cur_mhz * 1e6 / FP64_OPS_PER_CELL / 1e9
        achieved = cells_a1 / (fwd / 1000.0) / 1e9 if fwd > 0 else None
        # HBM side of the same kernel, algorithmic: rows read once (2 B per base), 4 bits written per DP cell, score + endrow
        alg_bytes = rows_bytes + cells_a1 // 2 + n * 12
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic, traffic_src = None, None
        try:       # dram__bytes_read + dram__bytes_write per alignment of this kernel from this round's ncu capture
            with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as fh:
                tj = json.load(fh)
            traffic = tj["a1_trace_bytes_per_alignment"] * n
            traffic_src = tj.get("source")
        except Exception:
            pass
        line = {
            "metric": "adaptorAlign reads/s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "gcups": gcups,
            "step_roofline_frac": gcups / world / peak_gcups,
            "config": {"workload": workload, "reads_per_gpu": n, "cells_per_read": cells_per_read,
                       "gapOpening": GO, "gapExtension": GE,
                       "l2": "inputs (%.0f MB packed windows per GPU) and traceback records exceed the 126 MB L2" % (2 * rows_bytes / 1e6)},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches,
            "roofline": {
                "bound": "fp64_alu", "kernel": dom_kernel, "achieved": achieved, "peak": peak_gcups, "unit": "GCUPS",
                "frac": (achieved / peak_gcups) if achieved else None,
                "peak_at_measured_clock": peak_gcups_at_clock,
                "frac_at_measured_clock": (achieved / peak_gcups_at_clock) if achieved else None,
                "how": "cells = n*250*70 per launch / CUDA-event duration of the forward launch; peak = %d SMs * 64 FP64 lanes * f / 10 FP64 ops per cell, f = clocks.max.sm" % sms,
                "launch_ms": fwd,
                "traffic": traffic, "traffic_source": traffic_src, "traffic_algorithmic": alg_bytes,
                "hbm": {"achieved": alg_bytes / (fwd / 1000.0) / 1e9 if fwd > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                        "frac": (alg_bytes / (fwd / 1000.0) / 1e9 / hbm_peak) if fwd > 0 else None,
                        "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
            },
            "cpu_baseline": cpu,
            "kernels": {"adaptor1": kernel_names[0], "adaptor2": kernel_names[1]},
        }
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def extra_records(args, ch, front, back, dptr, dout, enc, rank, world, local_rank, cores, peak_gcups, barrier, max_over_ranks):
    """Sub-records c3 (getAdaptorThresholds), c4 (barcodeAlign) and c5 (the north-star job)."""
    import torch
    from sarlacc_b200 import native, synth, _lib
    n = args.reads
    s1, e1 = setup_subseqs(A1)
    s2, e2 = setup_subseqs(A2)
    tdev = torch.device("cuda", local_rank)
    est = torch.cuda.ExternalStream(ch.stream(), device=tdev)
    out = {}

    # ---- c3: getAdaptorThresholds on the same reads: scramble both windows, four score-only passes, strand resolution
    # (R/getAdaptorThresholds.R:105-128), then the threshold selection over real and scrambled scores (:94-103)
    scr1 = torch.empty(n, dtype=torch.float64, device=tdev)
    scr2 = torch.empty(n, dtype=torch.float64, device=tdev)

    def c3_step():
        ch.scrambled_scores(GO, GE, A1, A2, seed=1, first_index=rank * n, score1=scr1.data_ptr(), score2=scr2.data_ptr())

    for _ in range(2):
        c3_step()
    ch.sync()
    ch.set_timing(True)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(2, args.steps)
    ev0.record(est)
    for _ in range(reps):
        c3_step()
    ch.join()
    ev1.record(est)
    ch.sync()
    ms_core = max_over_ranks(ev0.elapsed_time(ev1) / reps)
    ph = ch.phase_ms()
    ch.set_timing(False)
    t0 = time.perf_counter()
    thr1 = native.compute_threshold((dout["score1"].data_ptr(), n), (scr1.data_ptr(), n), 0.01, device=local_rank)
    thr2 = native.compute_threshold((dout["score2"].data_ptr(), n), (scr2.data_ptr(), n), 0.01, device=local_rank)
    ms_sel = max_over_ranks((time.perf_counter() - t0) * 1e3)
    cells = 2 * TOL * (len(A1) + len(A2))
    c3 = {"workload": "configs[2]: getAdaptorThresholds on %d reads per GPU (windows resident): device scramble of both windows, 4 score-only passes, "
                      "strand resolution, threshold selection (2 sorts + FDR scan per adaptor)" % n,
          "ms_core": ms_core, "ms_threshold_selection": ms_sel, "reads_per_s": world * n / ((ms_core + ms_sel) / 1e3),
          "gcups_core": world * n * cells / (ms_core / 1e3) / 1e9, "roofline_frac_core": n * cells / (ms_core / 1e3) / 1e9 / peak_gcups,
          "phases_ms_per_step": {"scramble": ph["scramble"] / reps, "score_only": ph["score_only"] / reps},
          "thresholds_rank0": [thr1, thr2]}
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = int(max(256, min(n, 5.0 * 0.02e9 * cores / cells)))
        kind, rate = cpu_score_only_rate(front, back, sample, cores)
        c3["cpu_baseline"] = {"value": rate, "unit": "reads/s", "cores": cores, "kind": kind,
                              "sample": "4 score-only alignments of the first %d reads' windows (unscrambled: same cells); the R scramble loop is not timed" % sample}
    out["c3"] = c3

    # ---- c4: barcodeAlign: --c4-sequences barcode-length sequences against 96 24-bp barcodes, one fused pass
    nb = args.c4_sequences
    barcodes = synth.random_barcodes(96, 24, 8, seed=3000)
    seqs, _ = synth.mock_barcode_sequences(nb, barcodes, seed=3000 + rank)
    native.barcode_align_multi(seqs, enc, GO, GE, barcodes)        # untimed: sizes the library's staging and device buffers
    dts = []
    for _ in range(3):
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bid, best, nxt = native.barcode_align_multi(seqs, enc, GO, GE, barcodes)
        torch.cuda.synchronize()
        dts.append(max_over_ranks(time.perf_counter() - t0))
    dt = statistics.median(dts)
    bc_cells = int(seqs.width().astype(np.int64).sum()) * 24 * 96
    c4 = {"workload": "configs[3]: barcodeAlign of %d sequences per GPU against 96 24-bp barcodes (host buffers in, best / next-best / id out; median of 3 calls)" % nb,
          "seconds_e2e": dt, "sequences_per_s": world * nb / dt, "gcups_e2e": world * bc_cells / dt / 1e9,
          "roofline_frac_e2e": bc_cells / dt / 1e9 / peak_gcups, "kernel": _lib.lib.sarlacc_version().decode()}
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle.oracle import Oracle, phred_encoding
        kind = "reference" if Oracle.available("ref") else "port"
        O = Oracle("ref" if kind == "reference" else "port")
        sample = int(max(64, min(nb, 5.0 * 0.02e9 * cores / (24 * 24 * 96))))
        sub = seqs[np.arange(sample)]
        t0 = time.perf_counter()
        for b in barcodes:
            O.align_score_only((sub.seq_pool, sub.seq_off), (sub.qual_pool, sub.qual_off), phred_encoding(), GO, GE, b, local=False, nthreads=cores)
        c4["cpu_baseline"] = {"value": sample / (time.perf_counter() - t0), "unit": "sequences/s", "cores": cores, "kind": kind,
                              "sample": "first %d sequences x 96 barcode_align calls; the R-level best / next-best passes are not timed" % sample}
    out["c4"] = c4

    # ---- c5: the north-star job: adaptorAlign + getAdaptorThresholds on --c5-reads reads, sharded by read index
    if args.c5_reads > 0:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import run_c5
        ch.sync()
        _lib.lib.sarlacc_trim_device_memory()
        rec = run_c5.run(args.c5_reads, 1136640, rank, world, local_rank, check_stride=max(1, args.c5_reads // world // 64) | 1)
        rec["roofline_frac_per_gpu"] = rec["gcups_per_gpu"] / peak_gcups
        rec["workload"] = "configs[4]: adaptorAlign + getAdaptorThresholds on %d synthetic 5 kb reads split over %d GPU(s) by read index: device " \
                          "read generation, 4 forward passes + kept-strand traceback, result columns to page-locked host tables, device scramble, " \
                          "4 score-only passes, scores gathered, thresholds selected on rank 0 -- wall time of all of it" % (args.c5_reads, world)
        out["c5"] = rec
    return out


if __name__ == "__main__":
    sys.exit(main())
