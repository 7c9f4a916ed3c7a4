"""Pins the CPU oracle (CPU-only tests):
  1. literal known answers of the reference's own tests and SURVEY 8(a)'s golden vectors,
  2. tests/golden/reference_vectors.json (generated from the reference C++ compiled verbatim),
  3. port (oracle/sarlacc_oracle.c) == ref (oracle/_ref, when built) bit for bit on random inputs.
"""
import json
import os

import numpy as np
import pytest

from conftest import stable_seed, VIGNETTE_A1, VIGNETTE_A2, random_windows

HERE = os.path.dirname(os.path.abspath(__file__))


def fromhex(xs):
    return np.array([float.fromhex(x) for x in xs], dtype=np.float64)


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "golden", "reference_vectors.json")) as fh:
        return json.load(fh)


def oracles(port, request):
    out = [port]
    from oracle.oracle import Oracle
    if Oracle.available("ref"):
        out.append(Oracle("ref"))
    return out


def test_survey_golden_vectors(port, enc, request):
    """adaptor AAAAGGGGCCCCTTTT, reads of tests/testthat/test-adaptor-align.R:7-19, Q20,
    go=5, ge=1, section = adaptor positions 5..8."""
    reads = ["AAAAGGGGCCCCTTTT", "ACGTACGTACGTAAAAGGGGCCCCTTTT", "AAAAGGGGCCCCTTTTACGTACGTACGT", "GGGGCCCCTTTT", "AAAAGGGGCCCC",
             "ACGTACGTACGTAAAAGGGGCCCCTTTTACGTACGTACGT", "ACGTACGTACGTAAAAGGGGCCCC", "GGGGCCCCTTTTACGTACGTACGT", "GGGGCCCC",
             "AAAAGGGGACGTCCCCTTTT", "AAAAGGCCTTTT"]
    expect = [(31.7680068848782, 1, 16, 5, 4), (31.7680068848782, 13, 28, 17, 4), (31.7680068848782, 1, 16, 5, 4),
              (14.8260051636586, 1, 12, 1, 4), (14.8260051636586, 1, 12, 5, 4), (31.7680068848782, 13, 28, 17, 4),
              (14.8260051636586, 13, 24, 17, 4), (14.8260051636586, 1, 12, 1, 4), (-2.11599655756092, 1, 8, 1, 4),
              (22.7680068848782, 1, 20, 5, 8), (14.8260051636586, 1, 12, 5, 2)]
    for O in oracles(port, request):
        sc, st, en, ss, sw = O.adaptor_align(reads, ["5" * len(r) for r in reads], enc, 5, 1, "AAAAGGGGCCCCTTTT", [4], [8])
        for i, (s, a, b, c, d) in enumerate(expect):
            assert sc[i] == pytest.approx(s, rel=1e-13)
            assert (st[i], en[i], ss[0][i], sw[0][i]) == (a, b, c, d)


def test_reference_test_known_answers(port, enc, request):
    for O in oracles(port, request):
        # empty adaptor -> 0, 0, 0 (test-adaptor-align.R:48-51)
        sc, st, en, _, _ = O.adaptor_align(["ACGTACGT"], ["55555555"], enc, 5, 1, "")
        assert (sc[0], st[0], en[0]) == (0.0, 0, 0)
        # empty read vs 16-mer -> -(16+5) (test-adaptor-align.R:53-56)
        sc, st, en, _, _ = O.adaptor_align([""], [""], enc, 5, 1, "AAAAGGGGCCCCTTTT")
        assert (sc[0], st[0], en[0]) == (-21.0, 0, 0)
        # Q10 cases (test-adaptor-align.R:67-84) and the UMI example of SURVEY 8(a)
        sc, st, en, _, _ = O.adaptor_align(["AAAAAAAAA"], ["+" * 9], enc, 5, 1, "AAACCCAAATTTAAA")
        assert sc[0] == pytest.approx(0.631972158994549, rel=1e-13) and (st[0], en[0]) == (1, 9)
        sc, st, en, _, _ = O.adaptor_align(["AAACCCAAA"], ["+" * 9], enc, 5, 1, "AAAAAA")
        assert sc[0] == pytest.approx(3.0879814393297, rel=1e-13) and (st[0], en[0]) == (1, 9)
        sc, st, en, ss, sw = O.adaptor_align(["AACGTAACGTACGTACGTGGGGGGG"], ["1234567890ABCDEFGHIJKLMNO"], enc, 5, 1, "AANNNAA", [2], [5])
        assert sc[0] == 7.9135845096322956 and (st[0], en[0], ss[0][0], sw[0][0]) == (1, 7, 3, 3)
        # section (1, nchar) returns the whole read including free end gaps (test-adaptor-align.R:120-121)
        sc, st, en, ss, sw = O.adaptor_align(["ACGTACGTACGTAAAAGGGGCCCCTTTTACGT"], ["5" * 32], enc, 5, 1, "AAAAGGGGCCCCTTTT", [0], [16])
        assert (ss[0][0], sw[0][0]) == (1, 32)
        # global mode
        g = O.align_score_only(["AAGGAATTAAGG", "AAGGATTAAGG", "AAGGAATTTAAGG", "GGCCAACCGGTT", ""], ["5" * 12, "5" * 11, "5" * 13, "5" * 12, ""],
                               enc, 4, 1, "AAGGAATTAAGG", local=False)
        np.testing.assert_allclose(g, [23.8260051636586, 16.8405047333537, 18.8260051636586, -16.0869974181707, -16.0], rtol=1e-13)
        m, mm, off = O.cost_tables(enc)
        assert off == b"!"
        assert (m[0][20], mm[0][20], mm[1][20], m[2][20], m[3][20]) == (
            1.9855004303048851, -6.2288186904958804, 0.99034982996628063, 0.41022048298106789, 0.0)


def test_golden_fixture(port, enc, golden, request):
    for O in oracles(port, request):
        for case in golden["adaptor_cases"]:
            sc, st, en, ss, sw = O.adaptor_align(case["seqs"], case["quals"], enc, case["go"], case["ge"], case["adaptor"],
                                                 case["sec_starts"], case["sec_ends"])
            assert np.array_equal(sc, fromhex(case["score"])), case["name"]
            assert st.tolist() == case["start"] and en.tolist() == case["end"], case["name"]
            assert ss.tolist() == case["sec_start"] and sw.tolist() == case["sec_width"], case["name"]
            g = O.align_score_only(case["seqs"], case["quals"], enc, case["go"], case["ge"], case["adaptor"], local=False)
            assert np.array_equal(g, fromhex(case["global_score"])), case["name"]
        gen = golden["general"]
        sc, ed, rs, qs = O.general_align(gen["seqs"], gen["quals"], enc, gen["go"], gen["ge"], gen["reference"])
        assert np.array_equal(sc, fromhex(gen["score"])) and ed.tolist() == gen["edit"]
        assert rs == gen["ref_aln"] and qs == gen["query_aln"]
        m, mm, off = O.cost_tables(enc)
        for r in range(4):
            assert np.array_equal(m[r], fromhex(golden["cost_tables"]["match"][r]))
            assert np.array_equal(mm[r], fromhex(golden["cost_tables"]["mismatch"][r]))


@pytest.mark.parametrize("adaptor,go,ge,alphabet,qlo,qhi", [
    (VIGNETTE_A1, 5, 1, "ACGT", 0, 50), (VIGNETTE_A2, 4, 1, "ACGT", 10, 40), ("ACACACNNNNACAC", 1, 1, "AC", 20, 20),
    ("ACGTNNRYACGTVVACKMBDHSW", 5, 1, "ACGTN", 0, 93), ("ACGTACGT", 0, 0, "ACGT", 5, 5), ("ACGTACGTAC", -1, 2, "ACGT", 0, 40),
    ("A", 10, 5, "ACGT", 0, 40), ("ACGT", 2.5, 0.3, "acgtACGT", 0, 40)])
def test_port_equals_reference_build(port, ref, enc, adaptor, go, ge, alphabet, qlo, qhi):
    rng = np.random.default_rng(stable_seed(adaptor, go))
    seqs, quals = random_windows(rng, 400, adaptor, 0, 90, qlo, qhi, alphabet=alphabet)
    import re
    st = [m.start() for m in re.finditer("[^ACTG]+", adaptor)]
    en = [m.end() for m in re.finditer("[^ACTG]+", adaptor)]
    a = port.adaptor_align(seqs, quals, enc, go, ge, adaptor, st, en, nthreads=3)
    b = ref.adaptor_align(seqs, quals, enc, go, ge, adaptor, st, en, nthreads=2)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert np.array_equal(port.align_score_only(seqs, quals, enc, go, ge, adaptor, local=False),
                          ref.align_score_only(seqs, quals, enc, go, ge, adaptor, local=False))
    ga, gb = port.general_align(seqs[:100], quals[:100], enc, go, ge, adaptor), ref.general_align(seqs[:100], quals[:100], enc, go, ge, adaptor)
    assert np.array_equal(ga[0], gb[0]) and np.array_equal(ga[1], gb[1]) and ga[2] == gb[2] and ga[3] == gb[3]


def test_error_messages(port, enc, request):
    from oracle.oracle import OracleError
    names, err = enc
    for O in oracles(port, request):
        with pytest.raises(OracleError, match="sequence and quality strings should have the same length"):
            O.adaptor_align(["ACGT"], ["555"], enc, 5, 1, "ACGT")
        with pytest.raises(OracleError, match="quality cannot be lower than smallest encoded value"):
            O.adaptor_align(["ACGT"], ["55 5"], enc, 5, 1, "ACGT")
        with pytest.raises(OracleError, match="unrecognized base in reference sequence"):
            O.adaptor_align(["ACGT"], ["5555"], enc, 5, 1, "ACXT")
        with pytest.raises(OracleError, match="encoding vector must be non-empty and named"):
            O.adaptor_align(["ACGT"], ["5555"], (None, err), 5, 1, "ACGT")
        with pytest.raises(OracleError, match="names of encoding vector must be one character in length"):
            O.adaptor_align(["ACGT"], ["5555"], (names[:3] + ["xy"] + names[4:], err), 5, 1, "ACGT")
        with pytest.raises(OracleError, match="names of encoding vector should increase consecutively"):
            O.adaptor_align(["ACGT"], ["5555"], (names[:3] + names[4:] + ["~"], err), 5, 1, "ACGT")
        with pytest.raises(OracleError, match="error probabilities should decrease"):
            O.adaptor_align(["ACGT"], ["5555"], (names, np.concatenate([err[:5], [1.0], err[6:]])), 5, 1, "ACGT")


def test_r_level_known_answers(port, enc):
    """tests/testthat/test-adaptor-align.R:125-139,186-206 and test-tuning.R:53-59, on the loop restatement."""
    from oracle import r_level as R
    assert R.setup_subseqs("AAAAGGNNNNCCTTTT") == ([7], [10])
    assert R.setup_subseqs("AAAAGGYYYYCCTTTT") == ([7], [10])
    assert R.setup_subseqs("AAAAGGNNNNCCRRRR") == ([7, 13], [10, 16])
    assert R.setup_subseqs("ACGT") == ([], [])
    f, b = R.get_front_and_back(["AAAACCCCGGGGTTTT", "ACG"], ["0123456789ABCDEF", "xyz"], 5)
    assert f == [("AAAAC", "01234"), ("ACG", "xyz")]
    assert b == [("AAAAC", "FEDCB"), ("CGT", "zyx")]
    read = "AACGTAACGTACGTACGTGGGGGGG"
    out = R.adaptor_align_R(port, enc, "AANNNAA", "CCCCCCC", [read, R.revcomp(read)], ["5" * 25, "5" * 25], go=5, ge=1)
    assert [o["reversed"] for o in out] == [False, True]
    for o in out:
        assert (o["adaptor1"]["start"], o["adaptor1"]["end"]) == (1, 7)
        assert (o["adaptor2"]["start"], o["adaptor2"]["end"]) == (25, 19)
        assert o["adaptor1"]["subseq"][0][0] == "CGT"
    assert out[0]["adaptor1"]["score"] == out[1]["adaptor1"]["score"]
    assert R.tied_overlap(range(1, 11), [x - 10 for x in range(1, 11)]) == 1
    assert R.tied_overlap(range(1, 11), range(1, 11)) == 0.5
    assert R.tied_overlap(range(1, 11), [x - 0.5 for x in range(1, 11)]) == pytest.approx(0.55)
    assert R.tied_overlap(range(1, 11), [x + 0.5 for x in range(1, 11)]) == pytest.approx(0.45)
    assert R.tied_overlap(range(1, 11), [x + 10 for x in range(1, 11)]) == 0
