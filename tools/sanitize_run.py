"""A small pass over every kernel of the hot path, for compute-sanitizer (memcheck / racecheck): the four-lane
and solo row-pair kernels with and without records, the one-row kernel, the literal kernel, traceback (both layouts and the
kept-strand form), packer, generator, scramble, strand resolution, threshold selection, UMI neighbours.  Results are
checked against the oracle so that a sanitizer-clean run is also a correct one.
usage: compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from sarlacc_b200 import native, synth  # noqa: E402
from oracle.oracle import Oracle, phred_encoding  # noqa: E402
from conftest import VIGNETTE_A1, VIGNETTE_A2, random_windows  # noqa: E402

O = Oracle("ref" if Oracle.available("ref") else "port")
enc = phred_encoding()
rng = np.random.default_rng(7)
S1, E1 = [16, 42], [28, 46]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300

front, back, widths, _ = synth.mock_windows(n, VIGNETTE_A1, VIGNETTE_A2, seed=11)
fa = (front.seq_pool, front.seq_off), (front.qual_pool, front.qual_off)
for adaptor, sec in ((VIGNETTE_A1, (S1, E1)), (VIGNETTE_A2, ([], []))):
    got = native.adaptor_align(front, enc, 5, 1, adaptor, *sec)
    exp = O.adaptor_align(*fa, enc, 5, 1, adaptor, *sec)
    assert all(np.array_equal(got[k], exp[k]) for k in range(3)), adaptor
    assert np.array_equal(native.adaptor_align_score_only(front, enc, 5, 1, adaptor), exp[0])
# ragged short reads: one-row kernel, masked steps, literal kernel (negative gap opening), general_align
seqs, quals = random_windows(rng, n, "ACGTNNRYACGTVVAC", 0, 60)
for go in (5, -1):
    got = native.adaptor_align((seqs, quals), enc, go, 2, "ACGTNNRYACGTVVAC", [4], [8])
    exp = O.adaptor_align(seqs, quals, enc, go, 2, "ACGTNNRYACGTVVAC", [4], [8])
    assert all(np.array_equal(got[k], exp[k]) for k in range(3))
g = native.general_align((seqs[:50], quals[:50]), enc, 4, 1, "AAGGAATTAAGGCCTTACGT")
e = O.general_align(seqs[:50], quals[:50], enc, 4, 1, "AAGGAATTAAGGCCTTACGT")
assert np.array_equal(g[0], e[0]) and list(g[2]) == list(e[2])
# barcodes: length-ordered fused pass
barcodes = synth.random_barcodes(8, 24, 8, seed=3)
bs, _ = synth.mock_barcode_sequences(n, barcodes, seed=4)
bid, best, nxt, mat = native.barcode_align_multi(bs, enc, 5, 1, barcodes, all_scores=True)
exp = np.stack([O.align_score_only((bs.seq_pool, bs.seq_off), (bs.qual_pool, bs.qual_off), enc, 5, 1, b, local=False) for b in barcodes])
assert np.array_equal(mat, exp)
# chunk engine: generator, both-ends pass with kept-strand traceback, scramble + score-only, thresholds
ch = native.Chunk(n, 250, enc)
ch.load_mock(n, VIGNETTE_A1, VIGNETTE_A2, seed=11)
w, rev, r1, r2 = ch.adaptor_align(5, 1, VIGNETTE_A1, VIGNETTE_A2, (S1, E1), ((), ()))
rev2, q1, q2 = native.adaptor_align_windows(front, back, enc, 5, 1, VIGNETTE_A1, VIGNETTE_A2, (S1, E1), ((), ()), read_width=widths)
assert np.array_equal(rev, rev2) and all(np.array_equal(r1[k], q1[k]) and np.array_equal(r2[k], q2[k]) for k in range(3))
s1, s2 = ch.scrambled_scores(5, 1, VIGNETTE_A1, VIGNETTE_A2, seed=3)
thr = native.compute_threshold(r1[0], s1, 0.01)
ch.close()
# speculative records (sub-ranges of 512+ reads): strand predictor, list-driven launches, and -- with the predictions
# inverted -- the re-run path; same results as the host-buffer entry on the same reads
ns = 576
os.environ["SARLACC_SPEC_MIN"] = "512"
f2, b2, w2, _ = synth.mock_windows(ns, VIGNETTE_A1, VIGNETTE_A2, seed=12)
ch = native.Chunk(ns, 250, enc)
ch.load_mock(ns, VIGNETTE_A1, VIGNETTE_A2, seed=12)
base = ch.adaptor_align(5, 1, VIGNETTE_A1, VIGNETTE_A2, (S1, E1), ((), ()))
os.environ["SARLACC_SPEC_TEST"] = "1"
flipped = ch.adaptor_align(5, 1, VIGNETTE_A1, VIGNETTE_A2, (S1, E1), ((), ()))
del os.environ["SARLACC_SPEC_TEST"]
fa2, ba2 = ((f2.seq_pool, f2.seq_off), (f2.qual_pool, f2.qual_off)), ((b2.seq_pool, b2.seq_off), (b2.qual_pool, b2.qual_off))
ea, eb = O.adaptor_align(*fa2, enc, 5, 1, VIGNETTE_A1, S1, E1), O.adaptor_align(*ba2, enc, 5, 1, VIGNETTE_A2)
ec, ed = O.adaptor_align(*ba2, enc, 5, 1, VIGNETTE_A1, S1, E1), O.adaptor_align(*fa2, enc, 5, 1, VIGNETTE_A2)
erev = (np.maximum(ea[0], 0) + np.maximum(eb[0], 0)) < (np.maximum(ec[0], 0) + np.maximum(ed[0], 0))
for got in (base, flipped):
    assert np.array_equal(got[1], erev)
    for k in range(3):
        assert np.array_equal(got[2][k], np.where(erev, ec[k], ea[k]))
    assert np.array_equal(got[3][0], np.where(erev, ed[0], eb[0]))
    assert np.array_equal(got[3][1], w2 - np.where(erev, ed[1], eb[1]) + 1)
ch.close()
# UMI neighbours
umis = ["".join(rng.choice(list("ACGT"), 12)) for _ in range(60)]
native.umi_group(umis + umis, 1)
print("sanitize_run ok: %d reads per case, threshold %.3f" % (n, thr))
