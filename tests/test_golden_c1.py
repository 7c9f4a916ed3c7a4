"""BASELINE.json configs[0] in full: 10 000 mockReads-style reads (~1.5 kb, vignette adaptors), all four alignments of
.align_AA_internal against the fixture tests/golden/c1_adaptor_align.npz, which tests/golden/make_c1_golden.py produced
from the reference's own reference_align.cpp (oracle/_ref).  CPU: the plain-C restatement reproduces the fixture.  GPU:
the library reproduces it -- the four reference calls one by one, the fused both-ends entry, and api.adaptorAlign."""
import os
import sys
import zlib

import numpy as np
import pytest

from conftest import VIGNETTE_A1, VIGNETTE_A2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "c1_adaptor_align.npz")
N, TOL = 10000, 250
S1, E1 = [16, 42], [28, 46]      # adaptor1's N runs: 0-based starts, 1-based ends (R/adaptorAlign.R:158)
KEYS = (("a1_front", "front", VIGNETTE_A1, True), ("a2_back", "back", VIGNETTE_A2, False),
        ("a1_back", "back", VIGNETTE_A1, True), ("a2_front", "front", VIGNETTE_A2, False))


@pytest.fixture(scope="module")
def c1():
    from sarlacc_b200 import synth
    from oracle import r_level as R
    g = np.load(GOLDEN)
    reads = synth.mock_reads(N, VIGNETTE_A1, VIGNETTE_A2, seed=1000)
    if zlib.crc32(reads.qual_pool.tobytes(), zlib.crc32(reads.seq_pool.tobytes())) != int(g["checksum"][0]):
        pytest.skip("numpy's Philox stream differs from the one the fixture was generated with")
    seqs, quals = reads.seq_strings(), reads.qual_strings()
    front, back = R.get_front_and_back(seqs, quals, TOL)
    win = {"front": ([w[0] for w in front], [w[1] for w in front]), "back": ([w[0] for w in back], [w[1] for w in back])}
    return g, reads, win


def check_against_golden(g, key, got, sections):
    assert np.array_equal(got[0], g[key + "_score"]), key
    assert np.array_equal(got[1], g[key + "_start"]) and np.array_equal(got[2], g[key + "_end"]), key
    if sections:
        for s in range(2):
            assert np.array_equal(got[3][s], g[key + "_sec_start"][s]) and np.array_equal(got[4][s], g[key + "_sec_width"][s]), (key, s)


def test_restatement_reproduces_the_c1_fixture(c1, port, enc):
    g, reads, win = c1
    for key, side, adaptor, sections in KEYS:
        got = port.adaptor_align(win[side][0], win[side][1], enc, 5, 1, adaptor, S1 if sections else [], E1 if sections else [],
                                 nthreads=os.cpu_count() or 1)
        check_against_golden(g, key, got, sections)
    assert np.mean(g["a1_front_score"]) > 25 and np.array_equal(g["width"], reads.width())


@pytest.mark.gpu
def test_library_reproduces_the_c1_fixture(c1, enc):
    from sarlacc_b200 import native, api, ReadSet
    g, reads, win = c1
    for key, side, adaptor, sections in KEYS:
        got = native.adaptor_align(win[side], enc, 5, 1, adaptor, S1 if sections else [], E1 if sections else [])
        check_against_golden(g, key, got, sections)
    # the fused entries keep, per read, the strand .resolve_strand picks (R/adaptorAlign.R:112-122,192-196) and flip adaptor2
    rev = (np.maximum(g["a1_front_score"], 0) + np.maximum(g["a2_back_score"], 0)) < (np.maximum(g["a1_back_score"], 0) + np.maximum(g["a2_front_score"], 0))
    width = g["width"].astype(np.int64)
    w, r, r1, r2 = native.adaptor_align_reads(reads, TOL, enc, 5, 1, VIGNETTE_A1, VIGNETTE_A2, (S1, E1), ((), ()))
    assert np.array_equal(w, width) and np.array_equal(r, rev) and 0.4 < rev.mean() < 0.6
    for k, name in enumerate(("score", "start", "end")):
        assert np.array_equal(r1[k], np.where(rev, g["a1_back_" + name], g["a1_front_" + name]))
    for s in range(2):
        assert np.array_equal(r1[3][s], np.where(rev, g["a1_back_sec_start"][s], g["a1_front_sec_start"][s]))
        assert np.array_equal(r1[4][s], np.where(rev, g["a1_back_sec_width"][s], g["a1_front_sec_width"][s]))
    assert np.array_equal(r2[0], np.where(rev, g["a2_front_score"], g["a2_back_score"]))
    assert np.array_equal(r2[1], width - np.where(rev, g["a2_front_start"], g["a2_back_start"]) + 1)
    assert np.array_equal(r2[2], width - np.where(rev, g["a2_front_end"], g["a2_back_end"]) + 1)
    ch = native.Chunk(N, TOL, enc)
    ch.load_reads(reads, TOL)
    w2, rr, q1, q2 = ch.adaptor_align(5, 1, VIGNETTE_A1, VIGNETTE_A2, (S1, E1), ((), ()))
    ch.close()
    assert np.array_equal(w2, width) and np.array_equal(rr, rev)
    for k in range(3):
        assert np.array_equal(q1[k], r1[k]) and np.array_equal(q2[k], r2[k])
    out = api.adaptorAlign(VIGNETTE_A1, VIGNETTE_A2, reads, number=3333)
    assert np.array_equal(out["reversed"], rev) and np.array_equal(out["adaptor1"]["score"], r1[0])
    assert np.array_equal(out["adaptor2"]["start"], r2[1]) and np.array_equal(out["adaptor1"]["end"], r1[2])
