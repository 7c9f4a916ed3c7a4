"""configs[3]: barcodeAlign of N barcode-length sequences against 96 24-bp barcodes, one fused pass from host buffers.
usage: python tools/bench_c4.py [n]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sarlacc_b200 import native, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
enc = native.phred_encoding()
barcodes = synth.random_barcodes(96, 24, 8, seed=3000)
seqs, _ = synth.mock_barcode_sequences(n, barcodes, seed=3000)
cells = int(seqs.width().astype(np.int64).sum()) * 24 * 96
native.barcode_align_multi(seqs[np.arange(min(n, 20000))], enc, 5, 1, barcodes)
best = None
for _ in range(3):
    t0 = time.perf_counter()
    native.barcode_align_multi(seqs, enc, 5, 1, barcodes)
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
print("barcodeAlign %d x 96: %.4f s = %.2f M sequences/s, %.0f GCUPS end to end (SOLO %s, length order %s)" %
      (n, best, n / best / 1e6, cells / best / 1e9, "off" if os.environ.get("SARLACC_NO_SOLO") else "on",
       "off" if os.environ.get("SARLACC_NO_LENGTH_ORDER") else "on"))
