"""Both-ends host-buffer call at small sizes, speculation on / off (which one is on comes from SARLACC_SPECULATE).
usage: python tools/ab/gpu_small.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sarlacc_b200 import native, synth
A1 = "ACGCAGATCGATCGATNNNNNNNNNNNNCGCGCGAGCTGACTNNNNGCACGACTCTGGTTTTTTTTTTTT"
A2 = "AAGGCCTTTTCCGACTCATGAA"
enc = native.phred_encoding()
for n in (2000, 10000, 28416, 56832, 113664):
    f, b, w, _ = synth.mock_windows_device(n, A1, A2, seed=9)
    keep = {}
    for _ in range(3):
        native.adaptor_align_windows(f, b, enc, 5, 1, A1, A2, ([16, 42], [28, 46]), ((), ()), read_width=w, reuse=keep)
    ts = []
    for _ in range(7):
        t0 = time.perf_counter()
        native.adaptor_align_windows(f, b, enc, 5, 1, A1, A2, ([16, 42], [28, 46]), ((), ()), read_width=w, reuse=keep)
        ts.append(time.perf_counter() - t0)
    print("n=%6d spec=%s: %.3f ms (median of 7)" % (n, os.environ.get("SARLACC_SPECULATE", "1"), 1e3 * float(np.median(ts))))
