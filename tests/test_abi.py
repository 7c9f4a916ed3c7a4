"""CPU-only: the C-ABI library loads, exports every symbol include/sarlacc_b200.h declares, validates arguments
with the reference's messages before touching a device, and fails loudly (no fallback) without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sarlacc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sarlacc_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from sarlacc_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(_lib.lib, s), "missing export: " + s
    assert set(_lib.EXPORTS) <= set(syms)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        getattr(raw, s)
    assert b"sm_100a" in _lib.lib.sarlacc_version()


def test_library_contains_sm100a_kernels():
    import subprocess
    from sarlacc_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def has_gpu():
    from sarlacc_b200 import _lib
    return _lib.lib.sarlacc_device_count() > 0


def test_validation_happens_before_the_device_is_touched():
    from sarlacc_b200 import native, SarlaccError
    names, err = native.phred_encoding()
    reads = (["ACGT"], ["5555"])
    with pytest.raises(SarlaccError, match="encoding vector must be non-empty and named"):
        native.adaptor_align(reads, (None, err), 5, 1, "ACGT")
    with pytest.raises(SarlaccError, match="names of encoding vector must be one character in length"):
        native.adaptor_align(reads, (names[:3] + ["xy"] + names[4:], err), 5, 1, "ACGT")
    with pytest.raises(SarlaccError, match="names of encoding vector should increase consecutively"):
        native.barcode_align(reads, (names[:3] + names[4:] + ["~"], err), 5, 1, "ACGT")
    with pytest.raises(SarlaccError, match="error probabilities should decrease"):
        native.adaptor_align_score_only(reads, (names, np.concatenate([err[:5], [1.0], err[6:]])), 5, 1, "ACGT")
    with pytest.raises(SarlaccError, match="section starts and ends should have the same length"):
        native.adaptor_align(reads, (names, err), 5, 1, "ACGT", [1, 2], [3])
    with pytest.raises(SarlaccError, match="gap opening penalty should be a numeric scalar"):
        native.adaptor_align(reads, (names, err), [5, 6], 1, "ACGT")
    with pytest.raises(SarlaccError, match="adaptor sequence should be a string"):
        native.adaptor_align(reads, (names, err), 5, 1, ["ACGT", "AC"])
    with pytest.raises(SarlaccError, match="sequence and quality vectors should have the same length"):
        native.adaptor_align((["ACGT", "A"], ["5555"]), (names, err), 5, 1, "ACGT")


def test_degenerate_reference_needs_no_device():
    """An empty adaptor runs no DP column (src/reference_align.cpp:82-90): score 0 / column-0 score, coordinates 0."""
    from sarlacc_b200 import native
    enc = native.phred_encoding()
    got = native.adaptor_align((["ACGT", ""], ["5555", ""]), enc, 5, 1, "", [], [])
    assert list(got[0]) == [0.0, 0.0] and list(got[1]) == [0, 0] and list(got[2]) == [0, 0]
    sc = native.barcode_align((["ACGT", "", "A"], ["5555", "", "5"]), enc, 5, 1, "")
    assert list(sc) == [-9.0, 0.0, -6.0]


def test_no_cpu_fallback():
    from sarlacc_b200 import native, SarlaccError
    if has_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(SarlaccError, match="requires a CUDA device"):
        native.adaptor_align((["ACGT"], ["5555"]), native.phred_encoding(), 5, 1, "ACGT")
    with pytest.raises(SarlaccError, match="requires a CUDA device"):
        native.Resident((["ACGT"], ["5555"]), native.phred_encoding())


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under sarlacc_b200/ may import, link or call it."""
    pkg = os.path.join(ROOT, "sarlacc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".cuh")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)


def test_no_predicate_spills_in_row_loops():
    """ptxas sometimes runs out of predicate registers in the unrolled row loop of a forward kernel and moves them through
    general registers (P2R + LOP3 + ISETP, or LOP3 alone: ten extra instructions per DP cell; sarlacc_b200/csrc/kernels.cu:
    SoloJit).  This watches the row loops of the geometries the vignette adaptors and 24-bp barcodes use."""
    import shutil
    import sys
    from sarlacc_b200 import _lib
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump unavailable")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_loop
    for C, solo in ((18, 0), (20, 1), (21, 1), (22, 1), (23, 1), (24, 1), (11, 0), (12, 0)):
        per_trace = sass_loop.analyse(_lib.LIB_PATH, "wf_forward2", C, 1, solo)[0]
        per_score = sass_loop.analyse(_lib.LIB_PATH, "wf_forward2", C, 0, solo)[0]
        assert per_trace < 31.0 and per_score < 25.0, (C, solo, per_trace, per_score)
