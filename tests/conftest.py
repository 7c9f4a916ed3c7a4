import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    from oracle.oracle import Oracle
    if not Oracle.available("port"):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return Oracle("port")


@pytest.fixture(scope="session")
def ref():
    from oracle.oracle import Oracle
    if not Oracle.available("ref"):
        pytest.skip("oracle/_ref/libsarlacc_ref.so not built (needs /root/reference)")
    return Oracle("ref")


@pytest.fixture(scope="session", params=["port", "ref"])
def both_oracles(request):
    """Every GPU parity test runs against the plain-C restatement AND the reference's own reference_align.cpp compiled
    verbatim (oracle/_ref): the GPU modules alias their `port` fixture to this one."""
    from oracle.oracle import Oracle
    kind = request.param
    if not Oracle.available(kind):
        if kind == "ref":
            pytest.skip("oracle/_ref/libsarlacc_ref.so not built (needs /root/reference)")
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return Oracle(kind)


def stable_seed(*parts):
    """A seed that is the same in every process (str hashes are salted per process; failing cases must replay)."""
    import zlib
    return zlib.crc32(repr(parts).encode())


@pytest.fixture(scope="session")
def enc():
    from oracle.oracle import phred_encoding
    return phred_encoding()


# vignettes/correction.Rmd:41-42
VIGNETTE_A1 = "ACGCAGATCGATCGATNNNNNNNNNNNNCGCGCGAGCTGACTNNNNGCACGACTCTGGTTTTTTTTTTTT"
VIGNETTE_A2 = "AAGGCCTTTTCCGACTCATGAA"


def random_windows(rng, n, adaptor, minlen=20, maxlen=80, qlo=0, qhi=40, embed=True, alphabet="ACGT"):
    """Random reads, most carrying a mutated copy of `adaptor` (N positions filled at random) at a random
    offset -- the generator of tests/testthat/test-adaptor-align.R:148-154 with our own RNG."""
    seqs, quals = [], []
    for _ in range(n):
        ln = int(rng.integers(minlen, maxlen + 1))
        body = rng.choice(list(alphabet), size=ln).tolist()
        if embed and adaptor and rng.random() < 0.8:
            ad = [c if c in "ACGT" else rng.choice(list("ACGT")) for c in adaptor]
            out = []
            for c in ad:   # substitutions / deletions / insertions
                r = rng.random()
                if r < 0.05:
                    out.append(rng.choice(list("ACGT")))
                elif r < 0.08:
                    continue
                elif r < 0.11:
                    out.extend([c, rng.choice(list("ACGT"))])
                else:
                    out.append(c)
            pos = int(rng.integers(0, max(1, ln - len(out) + 1)))
            body[pos:pos + len(out)] = out
            body = body[:ln] if len(body) > ln else body
        seqs.append("".join(body))
        quals.append("".join(chr(33 + int(q)) for q in rng.integers(qlo, qhi + 1, size=len(body))))
    return seqs, quals
