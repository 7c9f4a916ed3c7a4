"""Host-buffer path breakdown: python tools/e2e_probe.py [nreads] [reps]   (set SARLACC_DEBUG_TIMING=1 for phase times)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from sarlacc_b200 import native, synth, _lib  # noqa: E402

A1 = "ACGCAGATCGATCGATNNNNNNNNNNNNCGCGCGAGCTGACTNNNNGCACGACTCTGGTTTTTTTTTTTT"
A2 = "AAGGCCTTTTCCGACTCATGAA"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
front, back, widths, _ = synth.mock_windows_device(n, A1, A2, seed=2000)
enc = native.phred_encoding()
s1, e1 = [16, 42], [28, 46]
w = widths.astype(np.int32)
print("host cores", os.cpu_count(), "threads env", os.environ.get("SARLACC_HOST_THREADS"))
if os.environ.get("SARLACC_HOST_THREADS"):
    _lib.lib.sarlacc_set_host_threads(int(os.environ["SARLACC_HOST_THREADS"]))
if os.environ.get("PROBE_PINNED"):
    import torch
    from sarlacc_b200 import ReadSet
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()   # noqa: E731
    front = ReadSet(pin(front.seq_pool), front.seq_off, pin(front.qual_pool), front.qual_off, front.names)
    back = ReadSet(pin(back.seq_pool), back.seq_off, pin(back.qual_pool), back.qual_off, back.names)
    print("inputs pinned")
for r in range(reps):
    t0 = time.perf_counter()
    native.adaptor_align_windows(front, back, enc, 5.0, 1.0, A1, A2, (s1, e1), ([], []), read_width=w)
    dt = time.perf_counter() - t0
    print("adaptor_align_windows: %.1f ms  %.2f M reads/s" % (dt * 1e3, n / dt / 1e6), flush=True)
