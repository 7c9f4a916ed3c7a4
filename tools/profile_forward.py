"""Small driver for ncu: packs N synthetic windows, runs the forward(+traceback) pass a few times.
usage: python tools/profile_forward.py [nreads] [adaptor: a1|a2] [mode: trace|score] [reps]"""
import os
import sys
import time

os.environ.setdefault("SARLACC_NO_OVERLAP", "1")   # time the forward kernel alone (no traceback of the previous sub-range beside it)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sarlacc_b200 import native, synth  # noqa: E402

A1 = "ACGCAGATCGATCGATNNNNNNNNNNNNCGCGCGAGCTGACTNNNNGCACGACTCTGGTTTTTTTTTTTT"
A2 = "AAGGCCTTTTCCGACTCATGAA"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
which = sys.argv[2] if len(sys.argv) > 2 else "a1"
mode = sys.argv[3] if len(sys.argv) > 3 else "trace"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ad = A1 if which == "a1" else A2
front, back, _, _ = synth.mock_windows_device(n, A1, A2, seed=2000)
enc = native.phred_encoding()
r = native.Resident(front, enc)
r.set_timing(True)
ss, se = ([16, 42], [28, 46]) if which == "a1" else ([], [])
m = r.MODE_TRACE_LOCAL if mode == "trace" else r.MODE_SCORE_LOCAL
for _ in range(reps):
    r.align(m, 5, 1, ad, ss, se)
    ms = r.forward_ms()
    cells = r.cells(len(ad))
    print("%s %s: forward %.3f ms, %.1f GCUPS (%s)" % (which, mode, ms, cells / ms / 1e6, r.last_kernel()))
