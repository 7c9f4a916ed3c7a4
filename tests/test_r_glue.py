"""The R glue (sarlacc_b200/csrc/r_glue.cpp: SEXP adaptor_align(SEXP x 8) ... cluster_umis_test) compiled against the
stand-in R / Biostrings headers of tests/rstub/ and driven through the SEXP layer -- there is no R in this image.

CPU: every error the glue can raise without a device travels through Rf_error's longjmp under AddressSanitizer with no
leak (the reference unwinds first too: BEGIN_RCPP / END_RCPP, /root/reference/src/adaptor_align.cpp:12,76); a canary entry
point that raises while a std::vector is alive proves the harness would notice.
GPU: the reference's per-read errors through the glue, and a battery of calls (character vectors and DNAStringSet byte
codes) whose R-shaped results equal the library's Python face on the same reads."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import VIGNETTE_A1, random_windows

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "rstub", "_build")


@pytest.fixture(scope="module")
def drivers():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "rstub")])
    return os.path.join(BUILD, "glue_driver"), os.path.join(BUILD, "glue_driver_asan")


def run(exe, *args, asan=False):
    env = dict(os.environ)
    if asan:
        env["ASAN_OPTIONS"] = "detect_leaks=1:protect_shadow_gap=0:exitcode=23"
        env["LSAN_OPTIONS"] = "exitcode=23"
    p = subprocess.run([exe] + list(args), capture_output=True, text=True, env=env, timeout=600)
    records = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    return p, records


def test_error_paths_unwind_before_the_r_error(drivers):
    _, asan = drivers
    p, rec = run(asan, "errors", asan=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "LeakSanitizer" not in p.stderr and "AddressSanitizer" not in p.stderr
    cases = {r["case"]: r for r in rec if "case" in r}
    assert len(cases) >= 16 and all(r["ok"] for r in cases.values()), cases
    assert rec[-1] == {"failures": 0}
    # the messages are the reference's (src/utils.cpp:5-31, src/adaptor_align.cpp:23-31, src/quality_encoding.cpp:5-32)
    assert cases["vector lengths differ"]["error"] == "sequence and quality vectors should have the same length"
    assert cases["encoding names not consecutive"]["error"] == "names of encoding vector should increase consecutively"


def test_the_harness_notices_an_error_raised_over_live_objects(drivers):
    _, asan = drivers
    p, rec = run(asan, "canary", asan=True)
    assert rec and rec[0]["case"] == "canary" and rec[0]["error"].startswith("leaky_entry")
    assert p.returncode == 23 and "LeakSanitizer: detected memory leaks" in p.stderr


@pytest.mark.gpu
def test_per_read_errors_through_the_glue(drivers):
    plain, _ = drivers
    p, rec = run(plain, "gpu_errors")
    assert p.returncode == 0, p.stdout + p.stderr
    cases = {r["case"]: r for r in rec if "case" in r}
    assert len(cases) == 4 and all(r["ok"] for r in cases.values()), cases


@pytest.mark.gpu
def test_glue_results_match_the_library(drivers, enc, tmp_path):
    from sarlacc_b200 import native
    plain, _ = drivers
    rng = np.random.default_rng(99)
    seqs, quals = random_windows(rng, 120, VIGNETTE_A1, 30, 260, 2, 40)
    path = tmp_path / "reads.txt"
    path.write_text("".join("%s %s\n" % (s, q) for s, q in zip(seqs, quals)))
    p, rec = run(plain, "parity", str(path))
    assert p.returncode == 0, p.stdout + p.stderr
    hexf = lambda xs: np.array([float.fromhex(x) for x in xs])      # noqa: E731
    aa = [r for r in rec if r["call"] == "adaptor_align"]
    exp = native.adaptor_align((seqs, quals), enc, 5, 1, VIGNETTE_A1, [16, 42], [28, 46])
    assert len(aa) == 2 and {r["s4"] for r in aa} == {0, 1}
    for r in aa:       # character vector and DNAStringSet (Biostrings byte codes) inputs
        assert np.array_equal(hexf(r["score"]), exp[0]) and r["start"] == exp[1].tolist() and r["end"] == exp[2].tolist()
        assert r["sec_start"] == [x.tolist() for x in exp[3]] and r["sec_width"] == [x.tolist() for x in exp[4]]
    by = {r["call"]: r for r in rec}
    assert np.array_equal(hexf(by["adaptor_align_score_only"]["score"]), native.adaptor_align_score_only((seqs, quals), enc, 4, 2, "AAGGCCTTTTCCGACTCATGAA"))
    assert np.array_equal(hexf(by["barcode_align"]["score"]), native.barcode_align((seqs, quals), enc, 5, 1, "AAGGCCTTTTCCGACTCATGAACC"))
    g = native.general_align((seqs[:40], quals[:40]), enc, 4, 1, "AAGGAATTAAGGCCTTACGT")
    assert np.array_equal(hexf(by["general_align"]["score"]), g[0]) and by["general_align"]["edit"] == g[1].tolist()
    assert by["general_align"]["ref"] == list(g[2]) and by["general_align"]["query"] == list(g[3])
    umis = [s[:12] if len(s) >= 12 else "ACGTACGTACGT" for s in seqs]
    groups = [np.arange(1, 121, 2, dtype=np.int32), np.arange(2, 121, 2, dtype=np.int32)]
    for got, grp in zip(by["umi_group"]["groups"], groups):
        want = native.umi_group(umis, 1, groups=[grp])
        assert got["clusters"] == [c.tolist() for c in want]
    # the optional fused routines
    bid, best, nxt = native.barcode_align_multi((seqs, quals), enc, 5, 1, ["AAGGCCTTTTCCGACTCATGAACC", "ACGTACGTACGTACGTACGTACGT", "TTGACCAGTTGACCAGTTGACCAG"])
    m = by["barcode_align_multi"]
    assert m["id"] == bid.tolist() and np.array_equal(hexf(m["best"]), best) and np.array_equal(hexf(m["next"]), nxt)
    from sarlacc_b200 import ReadSet
    width, rev, r1, r2 = native.adaptor_align_reads(ReadSet.from_strings(seqs, quals), 100, enc, 5, 1, VIGNETTE_A1, "AAGGCCTTTTCCGACTCATGAA", ([16, 42], [28, 46]), ((), ()))
    g = by["adaptor_align_reads"]
    assert g["reversed"] == rev.astype(int).tolist() and g["width"] == width.tolist()
    for got, want in ((g["adaptor1"], r1), (g["adaptor2"], r2)):
        assert np.array_equal(hexf(got["score"]), want[0]) and got["start"] == want[1].tolist() and got["end"] == want[2].tolist()
        assert got["sec_start"] == [x.tolist() for x in want[3]] and got["sec_width"] == [x.tolist() for x in want[4]]
