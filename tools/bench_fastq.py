"""adaptorAlign from a FASTQ file: ingest alone (sequential whole-read reader vs parallel condensed reader)
and api.adaptorAlign end to end.  usage: python tools/bench_fastq.py [reads] [read_length]"""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sarlacc_b200 import api, read_fastq, read_fastq_condensed  # noqa: E402

A1 = "ACGCAGATCGATCGATNNNNNNNNNNNNCGCGCGAGCTGACTNNNNGCACGACTCTGGTTTTTTTTTTTT"
A2 = "AAGGCCTTTTCCGACTCATGAA"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
rng = np.random.default_rng(6000)
path = os.path.join(tempfile.gettempdir(), "sarlacc_bench_%d.fastq" % os.getpid())
a1 = np.frombuffer(A1.replace("N", "A").encode(), np.uint8)
a2rc = np.frombuffer(A2.encode()[::-1].translate(bytes.maketrans(b"ACGT", b"TGCA")), np.uint8)
with open(path, "wb") as fh:
    for lo in range(0, n, 10000):
        m = min(10000, n - lo)
        seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, (m, L))]
        seq[:, :len(a1)] = a1
        seq[:, L - len(a2rc):] = a2rc
        qual = rng.integers(45, 75, (m, L)).astype(np.uint8)
        for i in range(m):
            fh.write(b"@read_%d\n" % (lo + i))
            fh.write(seq[i].tobytes())
            fh.write(b"\n+\n")
            fh.write(qual[i].tobytes())
            fh.write(b"\n")
size = os.path.getsize(path)
try:
    t0 = time.perf_counter()
    k = sum(len(c) for c in read_fastq(path, 100000))
    dt = time.perf_counter() - t0
    print("ingest, sequential whole reads: %d reads, %.2f GB in %.2f s = %.0f k reads/s (%.2f GB/s)" % (k, size / 1e9, dt, k / dt / 1e3, size / dt / 1e9))
    for rep in range(2):
        t0 = time.perf_counter()
        k = sum(len(c) for c, _ in read_fastq_condensed(path, 250, 100000))
        dt = time.perf_counter() - t0
        print("ingest, parallel condensed (first/last 250 bases): %.2f s = %.0f k reads/s (%.2f GB/s)" % (dt, k / dt / 1e3, size / dt / 1e9))
    for rep in range(2):
        t0 = time.perf_counter()
        out = api.adaptorAlign(A1, A2, path, number=100000)
        dt = time.perf_counter() - t0
        print("api.adaptorAlign(path): %d reads in %.2f s = %.0f k reads/s; median adaptor1 score %.1f, %.1f%% reversed"
              % (len(out["reversed"]), dt, n / dt / 1e3, float(np.median(out["adaptor1"]["score"])), 100 * float(np.mean(out["reversed"]))))
    # the same file gzip-compressed (ShortRead reads .gz transparently): a tenth of the reads, inflated by zlib in the library
    import gzip
    gzpath = path + ".gz"
    m = max(1, n // 10)
    with open(path, "rb") as src, gzip.open(gzpath, "wb", compresslevel=1) as dst:
        for _ in range(4 * m):
            dst.write(src.readline())
    gsize = os.path.getsize(gzpath)
    t0 = time.perf_counter()
    k = sum(len(c) for c, _ in read_fastq_condensed(gzpath, 250, 100000))
    dt = time.perf_counter() - t0
    print("ingest, gzip (%.2f GB compressed), condensed: %d reads in %.2f s = %.0f k reads/s (%.2f GB/s inflated)"
          % (gsize / 1e9, k, dt, k / dt / 1e3, size * (m / n) / dt / 1e9))
    os.remove(gzpath)
finally:
    os.remove(path)
