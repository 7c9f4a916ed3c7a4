#!/usr/bin/env python
"""bench.py -- adaptorAlign throughput on B200 (BASELINE.json configs[1]) vs the reference's CPU path.

One "step" = adaptorAlign's hot path over one batch of synthetic mockReads-style reads: the four
local-global alignments with traceback that .align_AA_internal performs per read
(adaptor1 x front, adaptor2 x back, adaptor1 x back, adaptor2 x front; R/adaptorAlign.R:186-189),
tolerance 250, vignette adaptors (70 bp with two N-runs, 22 bp), go=5, ge=1  ->  46 000 DP cells per read.

  value   : reads/s with the packed windows already resident in HBM (CUDA events on the launching stream)
  e2e     : the same work through the host-buffer C-ABI calls (pack -> pinned -> H2D -> kernels -> D2H) + strand
            resolution, i.e. what the R wrapper would see
  roofline: the dominant kernel (wavefront forward pass for adaptor1) against the FP64 ALU roofline
            SMs * 64 lanes * f / 10 FP64 ops per cell, plus its HBM side for completeness
  cpu_baseline / --impl reference: the reference's own reference_align.cpp (oracle/_ref, compiled verbatim)
            on all host cores, on a bounded sample of the same reads.

Launch: `python bench.py --gpus N --steps K --warmup W`, or under torchrun for N > 1 (one rank per GPU,
reads sharded by index, no collective on the data path; weak scaling: every rank aligns --reads reads).
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


def resolve_strand(s1, s2, r1, r2):
    # R/adaptorAlign.R:112-122
    f = np.maximum(s1, 0) + np.maximum(s2, 0)
    r = np.maximum(r1, 0) + np.maximum(r2, 0)
    return f < r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=1_000_000, help="reads per GPU per step (configs[1]: 1M x 5 kb)")
    ap.add_argument("--e2e-reads", type=int, default=0, help="reads per step for the host-buffer leg (0 = same as --reads)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target duration of the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
        front, back, _, _ = synth.mock_windows(per_step, A1, A2, tolerance=TOL, seed=2000)
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
    from sarlacc_b200 import native, _lib

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

    n = args.reads
    enc = native.phred_encoding()
    s1, e1 = setup_subseqs(A1)
    s2, e2 = setup_subseqs(A2)
    # every rank generates its own shard of one big read set (read index keyed RNG)
    front, back, widths, _ = synth.mock_windows(n, A1, A2, tolerance=TOL, seed=2000, first_index=rank * n)
    dev = local_rank
    rf = native.Resident(front, enc, device=dev)
    rb = native.Resident(back, enc, device=dev)
    # a dedicated (non-null) stream: the library launches on it, and the timing events are recorded on it
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream, "need a non-null stream handle"
    T = native.Resident.MODE_TRACE_LOCAL

    def step(timing=False):
        rf.set_timing(timing)
        rf.align(T, GO, GE, A1, s1, e1, stream=stream)
        fwd = rf.forward_ms() if timing else 0.0
        rb.align(T, GO, GE, A2, s2, e2, stream=stream)
        rb.align(T, GO, GE, A1, s1, e1, stream=stream)
        rf.set_timing(False)
        rf.align(T, GO, GE, A2, s2, e2, stream=stream)
        return fwd

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    _lib.lib.sarlacc_kernel_launches(1)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    launches = int(_lib.lib.sarlacc_kernel_launches(0))
    ms_step = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
    clocks = sampler.stop() if sampler else None
    kernel_name = rf.last_kernel()

    # dominant kernel, timed alone with events inside the library (same stream), outside the timed region so that
    # the event synchronisation does not serialise it
    fwd_ms = []
    os.environ["SARLACC_NO_OVERLAP"] = "1"     # time the kernel alone: no traceback of the previous sub-range beside it
    for _ in range(max(3, args.steps)):
        fwd_ms.append(step(timing=True))
    torch.cuda.synchronize()
    del os.environ["SARLACC_NO_OVERLAP"]
    rf.align(T, GO, GE, A1, s1, e1, stream=stream)
    dom_kernel = rf.last_kernel()
    fwd = statistics.median(fwd_ms)
    cells_a1 = rf.cells(len(A1))

    value = world * n / (ms_step / 1000.0)
    gcups = value * cells_per_read / 1e9

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
            from sarlacc_b200 import ReadSet
            pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()   # noqa: E731
            return ReadSet(pin(rs.seq_pool), rs.seq_off, pin(rs.qual_pool), rs.qual_off, rs.names)

        inputs_pinned = True
        try:
            sub_f, sub_b = pinned(sub_f), pinned(sub_b)
        except RuntimeError:          # page-locking refused (memlock limit): the library stages the bytes itself
            inputs_pinned = False

        sub_w = widths[:ne].astype(np.int32)

        def e2e_step():
            # .align_AA_internal + adaptor2 flip in one C-ABI call: windows packed and uploaded once, four alignments
            # with traceback, strand resolution and row selection on the device, selected rows copied back
            return native.adaptor_align_windows(sub_f, sub_b, enc, GO, GE, A1, A2, (s1, e1), (s2, e2), read_width=sub_w)

        def e2e_step_unfused():
            a = native.adaptor_align(sub_f, enc, GO, GE, A1, s1, e1)
            b = native.adaptor_align(sub_b, enc, GO, GE, A2, s2, e2)
            c = native.adaptor_align(sub_b, enc, GO, GE, A1, s1, e1)
            d = native.adaptor_align(sub_f, enc, GO, GE, A2, s2, e2)
            return resolve_strand(a[0], b[0], c[0], d[0])

        e2e_step()
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = max(2, min(args.steps, 3))
        for _ in range(reps):
            e2e_step()
        torch.cuda.synchronize()
        dt = max_over_ranks((time.perf_counter() - t0) / reps)
        # raw bases + qualities of both windows, two 8-byte offsets and a 4-byte length per window, read widths
        h2d = int(2 * (sub_f.seq_off[-1] + sub_b.seq_off[-1]) + 2 * ne * (16 + 4) + ne * 4)
        d2h = ne * (1 + (8 + 4 + 4 + 8 * len(s1)) + (8 + 4 + 4 + 8 * len(s2)))
        t0 = time.perf_counter()
        e2e_step_unfused()
        torch.cuda.synchronize()
        dt_unfused = max_over_ranks(time.perf_counter() - t0)
        # the same call with ordinary (pageable) host buffers: the library gathers the bytes into its own pinned staging
        native.adaptor_align_windows(pageable_f, pageable_b, enc, GO, GE, A1, A2, (s1, e1), (s2, e2), read_width=sub_w)
        barrier()
        t0 = time.perf_counter()
        native.adaptor_align_windows(pageable_f, pageable_b, enc, GO, GE, A1, A2, (s1, e1), (s2, e2), read_width=sub_w)
        torch.cuda.synchronize()
        dt_pageable = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * ne / dt, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "reads_per_step": ne, "ms_per_step": dt * 1000.0,
               "path": "sarlacc_adaptor_align_windows: pinned host CSR buffers -> H2D of the raw bytes -> device packer -> 4 alignments + "
                       "traceback + strand resolution/selection on device -> D2H of the kept rows",
               "inputs_pinned": inputs_pinned, "pageable_inputs_reads_per_s": world * ne / dt_pageable, "host_threads_per_rank": max(1, cores // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))),
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

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        props = torch.cuda.get_device_properties(dev)
        sms = props.multi_processor_count
        max_mhz = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
        cur_mhz = (clocks or {}).get("sm_mhz") or max_mhz
        peak_gcups = sms * FP64_LANES_PER_SM * max_mhz * 1e6 / FP64_OPS_PER_CELL / 1e9
        peak_gcups_at_clock = sms * FP64_LANES_PER_SM * cur_mhz * 1e6 / FP64_OPS_PER_CELL / 1e9
        achieved = cells_a1 / (fwd / 1000.0) / 1e9 if fwd > 0 else None
        # HBM side of the same kernel: rows read once (2 B/base) + 4-bit records written (8-byte word per lane-row)
        alg_bytes = rf.nbytes() + n * (TOL + 8) * 4 * 16 + n * 12   # rows read once; 4 lanes x 16-byte trace word per row slot; score + endrow
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        line = {
            "metric": "adaptorAlign reads/s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "gcups": gcups,
            "config": {"workload": workload, "reads_per_gpu": n, "cells_per_read": cells_per_read,
                       "gapOpening": GO, "gapExtension": GE,
                       "l2": "inputs (%.0f MB packed windows per GPU) and traceback records exceed the 126 MB L2" % (2 * rf.nbytes() / 1e6)},
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
                # dram__bytes_read+write of this kernel per alignment in profiles/r01_ncu_summary_v4.txt (100 k alignments:
                # 0.082 GB read + 1.568 GB written) scaled to the alignments of one bench pass
                "traffic": 16503.0 * n, "traffic_algorithmic": alg_bytes,
                "hbm": {"achieved": alg_bytes / (fwd / 1000.0) / 1e9 if fwd > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                        "frac": (alg_bytes / (fwd / 1000.0) / 1e9 / hbm_peak) if fwd > 0 else None,
                        "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
            },
            "cpu_baseline": cpu,
            "kernels": {"step": kernel_name},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
