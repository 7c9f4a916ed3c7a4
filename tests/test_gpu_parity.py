"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit for bit.

Bar (BASELINE.json north_star): coordinates / section starts+widths identical; scores within 1e-5
relative -- the FP64 kernels are expected to be, and are asserted to be, bit-identical.
"""
import numpy as np
import pytest

from conftest import VIGNETTE_A1, VIGNETTE_A2, random_windows, stable_seed

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def port(both_oracles):
    """Both CPU oracles in turn: the restatement and the reference's own C++ (tests/conftest.py: both_oracles)."""
    return both_oracles


def setup_subseqs(adaptor):
    import re
    st, en = [], []
    for m in re.finditer("[^ACTG]+", adaptor):
        st.append(m.start() + 1)
        en.append(m.end())
    return st, en


def check_adaptor(native, oracle, enc, seqs, quals, adaptor, go, ge, sections=None, **kw):
    st, en = setup_subseqs(adaptor) if sections is None else sections
    ss = [x - 1 for x in st]
    got = native.adaptor_align((seqs, quals), enc, go, ge, adaptor, ss, en, **kw)
    exp = oracle.adaptor_align(seqs, quals, enc, go, ge, adaptor, ss, en)
    assert np.array_equal(got[0], exp[0]), "scores differ: max |d| = %g" % np.max(np.abs(got[0] - exp[0]))
    assert np.array_equal(got[1], exp[1])
    assert np.array_equal(got[2], exp[2])
    for s in range(len(ss)):
        assert np.array_equal(got[3][s], exp[3][s])
        assert np.array_equal(got[4][s], exp[4][s])
    sc = native.adaptor_align_score_only((seqs, quals), enc, go, ge, adaptor, **kw)
    assert np.array_equal(sc, exp[0])
    return got


def test_known_answers(port, enc):
    """SURVEY 8(a) golden vectors + tests/testthat/test-adaptor-align.R:48-56."""
    from sarlacc_b200 import native
    reads = ["AAAAGGGGCCCCTTTT", "ACGTACGTACGTAAAAGGGGCCCCTTTT", "GGGGCCCC", "AAAAGGGGACGTCCCCTTTT", "AAAAGGCCTTTT", ""]
    quals = ["5" * len(r) for r in reads]
    got = check_adaptor(native, port, enc, reads, quals, "AAAAGGGGCCCCTTTT", 5, 1, sections=([5], [8]))
    assert got[0][0] == pytest.approx(31.7680068848782, rel=1e-13)
    assert got[0][5] == -21.0 and got[1][5] == 0 and got[2][5] == 0
    assert list(got[1][:5]) == [1, 13, 1, 1, 1] and list(got[2][:5]) == [16, 28, 8, 20, 12]
    assert list(got[3][0][:5]) == [5, 17, 1, 5, 5] and list(got[4][0][:5]) == [4, 4, 4, 8, 2]
    got = check_adaptor(native, port, enc, ["AACGTAACGTACGTACGTGGGGGGG"], ["1234567890ABCDEFGHIJKLMNO"], "AANNNAA", 5, 1)
    assert got[0][0] == 7.9135845096322956 and got[3][0][0] == 3 and got[4][0][0] == 3


@pytest.mark.parametrize("adaptor", [VIGNETTE_A1, VIGNETTE_A2, "ACGT", "A", "AANNNAA", "ACGTNNNNACGTYYYYACGT"[:12] + "ACGT",
                                     "ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACG" * 3, "YYRRACGTACGT"])
@pytest.mark.parametrize("go,ge", [(5, 1), (4, 1), (10, 5), (0, 1), (2.5, 0.3)])
def test_adaptor_align_random(port, enc, adaptor, go, ge):
    from sarlacc_b200 import native
    rng = np.random.default_rng(stable_seed(adaptor, go, ge))
    seqs, quals = random_windows(rng, 300, adaptor, 1, 120)
    seqs += ["", "A", "ACGT" * 70]
    quals += ["", "I", "5" * 280]
    check_adaptor(native, port, enc, seqs, quals, adaptor, go, ge)


def test_unusual_penalties_and_encodings(port, enc):
    """Zero and negative extension penalties stay on the wavefront kernel (it only needs go >= 0); Phred 0 gives a -Inf
    match cost; error probabilities above 1 give NaN costs and take the literal kernel.  All must equal the oracle."""
    from sarlacc_b200 import native
    rng = np.random.default_rng(101)
    seqs, quals = random_windows(rng, 200, VIGNETTE_A2, 1, 90, qlo=0, qhi=3)      # lots of '!' (Q0)
    for go, ge in [(0, 0), (3, -0.25), (0, 2), (1e-3, 1e-3), (50, 0)]:
        check_adaptor(native, port, enc, seqs, quals, VIGNETTE_A2, go, ge)
        got = native.barcode_align((seqs, quals), enc, go, ge, VIGNETTE_A2)
        assert np.array_equal(got, port.align_score_only(seqs, quals, enc, go, ge, VIGNETTE_A2, local=False))
    names, err = enc
    weird = (names, np.concatenate([[3.0, 2.0], err[2:]]))                         # log of a negative number -> NaN
    seqs, quals = random_windows(rng, 100, VIGNETTE_A2, 1, 60, qlo=0, qhi=10)
    got = native.adaptor_align((seqs, quals), weird, 5, 1, VIGNETTE_A2)
    exp = port.adaptor_align(seqs, quals, weird, 5, 1, VIGNETTE_A2)
    assert np.array_equal(got[0], exp[0], equal_nan=True) and np.array_equal(got[1], exp[1]) and np.array_equal(got[2], exp[2])


def test_adaptor_align_ties(port, enc):
    """Uniform qualities and repetitive sequence: co-optimal paths everywhere, tie-breaking must match."""
    from sarlacc_b200 import native
    rng = np.random.default_rng(7)
    seqs = ["".join(rng.choice(list("AC"), size=int(rng.integers(30, 90)))) for _ in range(400)]
    quals = ["5" * len(s) for s in seqs]
    for adaptor in ["ACACACACNNNNNNACACAC", "AAAAAAAAAACCCCCCCCCC", VIGNETTE_A1]:
        check_adaptor(native, port, enc, seqs, quals, adaptor, 5, 1)
        check_adaptor(native, port, enc, seqs, quals, adaptor, 1, 1)


def test_windows_250(port, enc):
    """The production shape: 250-base windows against the vignette adaptors (both kernel geometries)."""
    from sarlacc_b200 import native
    rng = np.random.default_rng(11)
    for adaptor in (VIGNETTE_A1, VIGNETTE_A2):
        seqs, quals = random_windows(rng, 600, adaptor, 200, 250, qlo=12, qhi=40)
        check_adaptor(native, port, enc, seqs, quals, adaptor, 5, 1)
        check_adaptor(native, port, enc, seqs, quals, adaptor, 5, 1, views=True)


def test_all_sections(port, enc):
    """Every (start,end) pair of a 15-mer, as tests/testthat/test-adaptor-align.R:96-118 does with combn(15,2)."""
    from sarlacc_b200 import native
    rng = np.random.default_rng(3)
    adaptor = "ACGTTGCAAGGCTCA"
    st, en = zip(*[(a, b) for a in range(1, 16) for b in range(a, 16)])
    seqs, quals = random_windows(rng, 200, adaptor, 10, 60)
    check_adaptor(native, port, enc, seqs, quals, adaptor, 5, 1, sections=(list(st), list(en)))


def test_generic_kernel_paths(port, enc, monkeypatch):
    """Shapes outside the wavefront envelope take the literal kernel: mixed IUPAC classes, negative gap
    opening, very long references; plus the wavefront shapes forced through it."""
    from sarlacc_b200 import native
    rng = np.random.default_rng(5)
    seqs, quals = random_windows(rng, 150, "ACGTNNRYACGTVVAC", 5, 70)
    check_adaptor(native, port, enc, seqs, quals, "ACGTNNRYACGTVVAC", 5, 1)
    check_adaptor(native, port, enc, seqs, quals, "ACGTACGTAC", -1, 2)
    long_ref = "".join(rng.choice(list("ACGT"), size=400))
    seqs2, quals2 = random_windows(rng, 40, long_ref, 300, 450)
    check_adaptor(native, port, enc, seqs2, quals2, long_ref, 5, 1)
    monkeypatch.setenv("SARLACC_FORCE_GENERIC", "1")
    check_adaptor(native, port, enc, seqs, quals, VIGNETTE_A1, 5, 1)


@pytest.mark.parametrize("force", ["1,12", "2,6", "4,3", "8,2", "16,1", "32,1", "4,12"])
def test_every_geometry(port, enc, monkeypatch, force):
    from sarlacc_b200 import native
    monkeypatch.setenv("SARLACC_FORCE_GC", force)
    rng = np.random.default_rng(17)
    adaptor = "ACGTTNNNNCAT"
    seqs, quals = random_windows(rng, 300, adaptor, 1, 100)
    check_adaptor(native, port, enc, seqs, quals, adaptor, 5, 1)


@pytest.mark.parametrize("force,L", [("4,18", 70), ("2,18", 35), ("2,16", 31), ("4,14", 55), ("1,18", 18), ("8,16", 125)])
def test_wide_geometries(port, enc, monkeypatch, force, L):
    """14-18 columns per lane (128-bit trace words beyond 16): used for score-only runs, forced here for both."""
    from sarlacc_b200 import native
    monkeypatch.setenv("SARLACC_FORCE_GC", force)
    rng = np.random.default_rng(L)
    adaptor = "".join(rng.choice(list("ACGTN"), size=L, p=[0.22, 0.22, 0.22, 0.22, 0.12]))
    seqs, quals = random_windows(rng, 250, adaptor, 1, 200)
    check_adaptor(native, port, enc, seqs, quals, adaptor, 5, 1)
    got = native.barcode_align((seqs, quals), enc, 4, 1, adaptor)
    assert np.array_equal(got, port.align_score_only(seqs, quals, enc, 4, 1, adaptor, local=False))


def test_barcode_align(port, enc):
    from sarlacc_b200 import native
    rng = np.random.default_rng(23)
    barcodes = ["".join(rng.choice(list("ACGT"), size=24)) for _ in range(12)]
    seqs, quals = [], []
    for i in range(500):
        b = list(barcodes[int(rng.integers(0, len(barcodes)))])
        for k in range(len(b)):
            if rng.random() < 0.08:
                b[k] = rng.choice(list("ACGT"))
        if rng.random() < 0.2:
            del b[int(rng.integers(0, len(b)))]
        if rng.random() < 0.2:
            b.insert(int(rng.integers(0, len(b))), rng.choice(list("ACGT")))
        seqs.append("".join(b))
        quals.append("".join(chr(33 + int(q)) for q in rng.integers(5, 40, size=len(b))))
    seqs.append("")
    quals.append("")
    exp = np.stack([port.align_score_only(seqs, quals, enc, 5, 1, b, local=False) for b in barcodes])
    for k, b in enumerate(barcodes[:3]):
        got = native.barcode_align((seqs, quals), enc, 5, 1, b)
        assert np.array_equal(got, exp[k])
    bid, best, nxt, mat = native.barcode_align_multi((seqs, quals), enc, 5, 1, barcodes, all_scores=True)
    assert np.array_equal(mat, exp)
    # R/barcodeAlign.R:12-35 running best / next best with strict >
    cur = np.full(len(seqs), -np.inf); nb = np.full(len(seqs), -np.inf); cid = np.zeros(len(seqs), np.int32)
    for k in range(len(barcodes)):
        s = exp[k]
        keep = s > cur
        second = ~keep & (s > nb)
        cid[keep] = k + 1
        nb[keep] = cur[keep]
        cur[keep] = s[keep]
        nb[second] = s[second]
    assert np.array_equal(bid, cid) and np.array_equal(best, cur) and np.array_equal(nxt, nb)
    bid2, best2, nxt2 = native.barcode_align_multi((seqs, quals), enc, 5, 1, barcodes)
    assert np.array_equal(bid2, cid) and np.array_equal(best2, cur) and np.array_equal(nxt2, nb)


def test_general_align(port, enc):
    from sarlacc_b200 import native
    rng = np.random.default_rng(29)
    ref = "AAGGAATTAAGGCCTTACGT"
    seqs, quals = random_windows(rng, 200, ref, 5, 40)
    seqs += ["", "AAGGATTAAGG", "AAGGAATTTAAGG"]
    quals += ["", "5" * 11, "5" * 13]
    for go, ge in [(4, 1), (5, 1), (10, 2)]:
        got = native.general_align((seqs, quals), enc, go, ge, ref)
        exp = port.general_align(seqs, quals, enc, go, ge, ref)
        assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1])
        assert got[2] == exp[2] and got[3] == exp[3]
        got = native.general_align((seqs, quals), enc, go, ge, ref, edit_only=True)
        assert np.array_equal(got[1], exp[1]) and got[2] == []


def test_errors(port, enc):
    from sarlacc_b200 import native, SarlaccError
    from oracle.oracle import OracleError
    cases = [
        (["ACGT", "ACGT"], ["5555", "555"], "ACGT"),          # length mismatch
        (["ACGT", "ACGT"], ["5555", "55 5"], "ACGT"),         # quality below '!'
        (["ACGT"], ["5555"], "ACXT"),                         # bad reference base, later column
        (["ACGT"], ["5 55"], "ACXT"),                         # ... bad quality wins
        (["ACGT"], ["5 55"], "XCGT"),                         # ... unless the first column is bad
        (["", "ACGT"], ["", "5555"], "acgt"),                 # lower-case adaptor (barcodeAlign quirk)
    ]
    for seqs, quals, adaptor in cases:
        with pytest.raises(OracleError) as e1:
            port.adaptor_align(seqs, quals, enc, 5, 1, adaptor)
        with pytest.raises(SarlaccError) as e2:
            native.adaptor_align((seqs, quals), enc, 5, 1, adaptor)
        assert str(e1.value) == str(e2.value)
    # empty reads never touch the reference
    got = native.adaptor_align((["", ""], ["", ""]), enc, 5, 1, "ACXT")
    assert np.array_equal(got[0], port.adaptor_align(["", ""], ["", ""], enc, 5, 1, "ACXT")[0]) and got[0][0] == -9.0
    # encoding errors
    names, err = enc
    for bad in [(None, err), (names[:5] + ["ab"] + names[6:], err), (names[:5] + names[6:] + ["~"], err),
                (names, np.concatenate([err[:10], [1.0], err[11:]]))]:
        with pytest.raises(OracleError) as e1:
            port.adaptor_align(["ACGT"], ["5555"], bad, 5, 1, "ACGT")
        with pytest.raises(SarlaccError) as e2:
            native.adaptor_align((["ACGT"], ["5555"]), bad, 5, 1, "ACGT")
        assert str(e1.value) == str(e2.value)


def test_degenerate(port, enc):
    from sarlacc_b200 import native
    got = native.adaptor_align((["ACGT", ""], ["5555", ""]), enc, 5, 1, "", [], [])
    assert list(got[0]) == [0.0, 0.0] and list(got[1]) == [0, 0]
    exp = port.align_score_only(["ACGT", "", "A"], ["5555", "", "5"], enc, 5, 1, "", local=False)
    assert np.array_equal(native.barcode_align((["ACGT", "", "A"], ["5555", "", "5"]), enc, 5, 1, ""), exp)
    got = native.adaptor_align(([], []), enc, 5, 1, "ACGT", [1], [2])
    assert len(got[0]) == 0 and len(got[3][0]) == 0


def test_quality_clamp_and_biostrings_codes(port, enc):
    """Qualities above the table are clamped (src/reference_align.cpp:218-221); Biostrings byte codes decode
    like DNAdecode (src/DNA_input.cpp:64-75)."""
    from sarlacc_b200 import native, SEQ_BIOSTRINGS
    names, err = enc
    short = (names[:40], err[:40])
    rng = np.random.default_rng(31)
    seqs, quals = random_windows(rng, 100, VIGNETTE_A2, 10, 60, qlo=0, qhi=60)
    check_adaptor(native, port, short, seqs, quals, VIGNETTE_A2, 5, 1)
    code = {"A": 1, "C": 2, "G": 4, "T": 8, "N": 15, "M": 3}
    seqs2 = [s[:5] + "N" + s[5:] + "M" for s in seqs]
    quals2 = [q[:5] + "5" + q[5:] + "5" for q in quals]
    coded = [bytes(code[c] for c in s) for s in seqs2]
    exp = port.adaptor_align(seqs2, quals2, short, 5, 1, VIGNETTE_A2)
    got = native.adaptor_align((coded, quals2), short, 5, 1, VIGNETTE_A2, seq_encoding=SEQ_BIOSTRINGS)
    assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1]) and np.array_equal(got[2], exp[2])


def test_resident(port, enc):
    from sarlacc_b200 import native
    rng = np.random.default_rng(37)
    seqs, quals = random_windows(rng, 500, VIGNETTE_A1, 150, 250, qlo=12, qhi=40)
    st, en = setup_subseqs(VIGNETTE_A1)
    ss = [x - 1 for x in st]
    exp = port.adaptor_align(seqs, quals, enc, 5, 1, VIGNETTE_A1, ss, en)
    r = native.Resident((seqs, quals), enc)
    assert r.cells(len(VIGNETTE_A1)) == sum(map(len, seqs)) * len(VIGNETTE_A1)
    r.align(r.MODE_TRACE_LOCAL, 5, 1, VIGNETTE_A1, ss, en)
    got = r.fetch()
    assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1]) and np.array_equal(got[2], exp[2])
    for s in range(len(ss)):
        assert np.array_equal(got[3][s], exp[3][s]) and np.array_equal(got[4][s], exp[4][s])
    r.align(r.MODE_SCORE_LOCAL, 5, 1, VIGNETTE_A2)
    assert np.array_equal(r.fetch(), port.align_score_only(seqs, quals, enc, 5, 1, VIGNETTE_A2))
    r.align(r.MODE_SCORE_GLOBAL, 4, 1, VIGNETTE_A2)
    assert np.array_equal(r.fetch(), port.align_score_only(seqs, quals, enc, 4, 1, VIGNETTE_A2, local=False))
    r.close()


def test_chunking_and_scratch_budget(port, enc, monkeypatch):
    """Results must not depend on chunk size or scratch budget (batching invariance,
    tests/testthat/test-general-align.R:81-93)."""
    from sarlacc_b200 import native
    rng = np.random.default_rng(41)
    seqs, quals = random_windows(rng, 700, VIGNETTE_A1, 50, 250)
    st, en = setup_subseqs(VIGNETTE_A1)
    ss = [x - 1 for x in st]
    exp = port.adaptor_align(seqs, quals, enc, 5, 1, VIGNETTE_A1, ss, en)
    monkeypatch.setenv("SARLACC_CHUNK", "97")
    monkeypatch.setenv("SARLACC_SCRATCH_MB", "16")
    check_adaptor(native, port, enc, seqs, quals, VIGNETTE_A1, 5, 1)
    r = native.Resident((seqs, quals), enc)
    r.align(r.MODE_TRACE_LOCAL, 5, 1, VIGNETTE_A1, ss, en)
    got = r.fetch()
    for s in range(len(ss)):
        assert np.array_equal(got[3][s], exp[3][s]) and np.array_equal(got[4][s], exp[4][s])
    assert np.array_equal(got[1], exp[1]) and np.array_equal(got[0], exp[0])


def test_c4_ninety_six_barcodes(port, enc):
    """BASELINE.json configs[3] in miniature: barcodeAlign of barcode-length sequences against 96 24-bp barcodes in one
    fused pass (sequences walked in order of length, one thread per alignment) == 96 barcode_align calls of the oracle +
    the running best / next-best of R/barcodeAlign.R:28-34."""
    from sarlacc_b200 import native, synth
    barcodes = synth.random_barcodes(96, 24, 8, seed=3000)
    seqs, pick = synth.mock_barcode_sequences(3000, barcodes, seed=3001)
    assert len(set(seqs.width().tolist())) > 4          # indels: a spread of lengths around 24
    bid, best, nxt, mat = native.barcode_align_multi(seqs, enc, 5, 1, barcodes, all_scores=True)
    arg = (seqs.seq_pool, seqs.seq_off), (seqs.qual_pool, seqs.qual_off)
    exp = np.stack([port.align_score_only(*arg, enc, 5, 1, b, local=False, nthreads=8) for b in barcodes])
    assert np.array_equal(mat, exp)
    e_best = np.full(len(seqs), -np.inf)
    e_next = np.full(len(seqs), -np.inf)
    e_id = np.zeros(len(seqs), np.int32)
    for b in range(96):
        keep = exp[b] > e_best
        second = ~keep & (exp[b] > e_next)
        e_next[keep] = e_best[keep]
        e_best[keep] = exp[b][keep]
        e_id[keep] = b + 1
        e_next[second] = exp[b][second]
    assert np.array_equal(bid, e_id) and np.array_equal(best, e_best) and np.array_equal(nxt, e_next)
    assert np.mean(bid - 1 == pick) > 0.95
    # one barcode at a time (barcode_align, the reference's own call) goes through the same length-ordered walk
    assert np.array_equal(native.barcode_align(seqs, enc, 5, 1, barcodes[5]), exp[5])


def test_barcode_sequences_beyond_one_round_of_the_grid(port, enc):
    """More barcode-length sequences than one launch's lanes (56 832 in the one-thread-per-alignment kernel): the chunk
    is walked longest first with every second round of lanes turned round, and chunks overlap on the device -- neither
    may show in the results."""
    from sarlacc_b200 import native, synth
    barcodes = synth.random_barcodes(3, 24, 8, seed=3100)
    seqs, _ = synth.mock_barcode_sequences(300000, barcodes, seed=3101)
    bid, best, nxt, mat = native.barcode_align_multi(seqs, enc, 5, 1, barcodes, all_scores=True)
    arg = (seqs.seq_pool, seqs.seq_off), (seqs.qual_pool, seqs.qual_off)
    exp = np.stack([port.align_score_only(*arg, enc, 5, 1, b, local=False, nthreads=8) for b in barcodes])
    assert np.array_equal(mat, exp)
    order = np.argsort(-exp, axis=0, kind="stable")
    assert np.array_equal(best, np.take_along_axis(exp, order[:1], 0)[0])
    assert np.array_equal(bid, order[0] + 1)            # ties: the first barcode wins (strict >)
    sec = np.take_along_axis(exp, order[1:2], 0)[0]
    assert np.array_equal(nxt, sec)
