"""UMI grouping timings: device neighbour search + host clustering (sarlacc_umi_group) vs the reference's
own trie search + clustering (oracle/_ref/libsarlacc_umi_ref.so, one thread, as the reference runs it).
usage: python tools/bench_umi.py [reads] [group_size]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sarlacc_b200 import native  # noqa: E402
from oracle.umi import UmiRef  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
gsize = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
rng = np.random.default_rng(4000)
L = 12
copies = 8
nm = n // copies
mol = rng.integers(0, 4, size=(nm, L), dtype=np.int8)
reads = np.repeat(mol, copies, axis=0)
sub = rng.random(reads.shape) < 0.03
reads[sub] = rng.integers(0, 4, size=int(sub.sum()))
reads = reads[rng.permutation(len(reads))]
pool = np.frombuffer(b"ACGT", np.uint8)[reads].reshape(-1).copy()
off = np.arange(0, (len(reads) + 1) * L, L, dtype=np.int64)
n = len(reads)
groups = [np.arange(a + 1, min(a + gsize, n) + 1, dtype=np.int32) for a in range(0, n, gsize)]
pairs = sum(len(g) ** 2 for g in groups)
native.umi_group((pool[:L * 100], off[:101]), 1)
for thr in (1, 3):
    t0 = time.perf_counter()
    cl = native.umi_group((pool, off), thr, groups=groups)
    dt = time.perf_counter() - t0
    print("umi_group threshold %d: %d reads in %d pre-groups of %d (%.2e pairs): %.3f s = %.2f M reads/s, %.1f G pairs/s; %d clusters"
          % (thr, n, len(groups), gsize, pairs, dt, n / dt / 1e6, pairs / dt / 1e9, len(cl)))
if UmiRef.available():
    R = UmiRef()
    m = min(n, 40 * gsize)
    seqs = [bytes(pool[off[i]:off[i + 1]]).decode() for i in range(m)]
    gs = [g.tolist() for g in groups[:m // gsize]]
    for thr in (1, 3):
        t0 = time.perf_counter()
        ref = R.umi_group(seqs, thr, None, None, gs)
        dt = time.perf_counter() - t0
        got = native.umi_group((pool[:off[m]], off[:m + 1]), thr, groups=gs)
        same = [c.tolist() for c in got] == ref
        print("reference umi_group threshold %d (1 thread): %d reads in %.3f s = %.3f M reads/s; identical clusters: %s" % (thr, m, dt, m / dt / 1e6, same))
