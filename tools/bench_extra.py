"""Secondary measurements quoted in DESIGN.md (not the bench.py contract): configs[2] getAdaptorThresholds-style
score-only passes on scrambled windows, configs[3] barcodeAlign of 1 M sequences x 96 24-bp barcodes (fused pass)
vs 96 separate barcode_align calls.  usage: python tools/bench_extra.py [n]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sarlacc_b200 import api, native, synth  # noqa: E402

A1 = "ACGCAGATCGATCGATNNNNNNNNNNNNCGCGCGAGCTGACTNNNNGCACGACTCTGGTTTTTTTTTTTT"
A2 = "AAGGCCTTTTCCGACTCATGAA"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
if os.environ.get("SARLACC_HOST_THREADS"):
    from sarlacc_b200 import _lib
    _lib.lib.sarlacc_set_host_threads(int(os.environ["SARLACC_HOST_THREADS"]))
enc = native.phred_encoding()

# ---- configs[3]: barcodeAlign --------------------------------------------------------------------------------
barcodes = synth.random_barcodes(96, 24, 8, seed=3000)
seqs, pick = synth.mock_barcode_sequences(n, barcodes, seed=3000)
cells = int(seqs.width().sum()) * 24 * 96
native.barcode_align_multi(seqs[np.arange(1000)], enc, 5, 1, barcodes)
t0 = time.perf_counter()
native.barcode_align_multi(seqs, enc, 5, 1, barcodes)
print("(first full-size call, staging buffers grow: %.3f s)" % (time.perf_counter() - t0))
t0 = time.perf_counter()
bid, best, nxt = native.barcode_align_multi(seqs, enc, 5, 1, barcodes)
dt = time.perf_counter() - t0
print("barcodeAlign fused: %d sequences x 96 barcodes in %.3f s = %.2f M seq/s, %.1f GCUPS end to end (host buffers); accuracy %.4f"
      % (n, dt, n / dt / 1e6, cells / dt / 1e9, float(np.mean(bid - 1 == pick))))
sub = seqs[np.arange(min(n, 200000))]
t0 = time.perf_counter()
for b in barcodes[:8]:
    native.barcode_align(sub, enc, 5, 1, b)
dt8 = time.perf_counter() - t0
print("barcode_align one call per barcode (reference call pattern): %.3f s for 8 barcodes x %d sequences -> %.2f M seq/s for 96"
      % (dt8, len(sub), len(sub) / (dt8 * 12) / 1e6))

# ---- configs[2]: thresholds (score-only on scrambled windows, windows resident) -----------------------------
m = min(n, 1000000)
front, back, widths, _ = synth.mock_windows_device(m, A1, A2, seed=2000)
rf0, rb0 = native.Resident(front, enc), native.Resident(back, enc)
t0 = time.perf_counter()
rf = rf0.scrambled(1, first_index=0, stream_id=0)
rb = rb0.scrambled(1, first_index=0, stream_id=1)
t_scr = time.perf_counter() - t0
for r, a in ((rf, A1), (rb, A2), (rb, A1), (rf, A2)):
    r.align(r.MODE_SCORE_LOCAL, 5, 1, a)
    r.fetch()
t0 = time.perf_counter()
sc = {}
for key, r, a in (("S", rf, A1), ("E", rb, A2), ("RS", rb, A1), ("RE", rf, A2)):
    r.align(r.MODE_SCORE_LOCAL, 5, 1, a)
    sc[key] = r.fetch()
dt = time.perf_counter() - t0
rev = api._resolve_strand(sc["S"], sc["E"], sc["RS"], sc["RE"])["reversed"]
print("getAdaptorThresholds core: %d reads, 4 score-only passes on scrambled windows in %.3f s = %.2f M reads/s, %.1f GCUPS (resident); "
      "device scramble of 2 x %d windows %.3f s; %.1f%% reversed" % (m, dt, m / dt / 1e6, m * 46000 / dt / 1e9, m, t_scr, 100 * rev.mean()))
