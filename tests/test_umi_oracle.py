"""CPU-only: pins the UMI-grouping oracles.  oracle/_ref/libsarlacc_umi_ref.so is the reference's own
umi_group / sorted_trie / cluster_umis compiled verbatim; oracle/umi.py:port_* is the restatement.  Checked here:
restatement == reference on seeded inputs, both == the committed golden vectors, and the reference tests' own slow R
check (tests/testthat/test-umicluster.R:4-29) restated."""
import json
import os

import numpy as np
import pytest

from umi_cases import CASES, make_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def attempt(f, *a):
    try:
        return {"value": [list(map(int, x)) for x in f(*a)]}
    except RuntimeError as e:
        return {"error": str(e)}


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "umi_vectors.json")) as fh:
        return json.load(fh)["cases"]


@pytest.fixture(scope="module")
def umiref():
    from oracle.umi import UmiRef
    if not UmiRef.available():
        pytest.skip("oracle/_ref/libsarlacc_umi_ref.so not built (needs /root/reference)")
    return UmiRef()


def test_golden_inputs_are_reproducible(golden):
    for c, (name, seed, kw, t1, t2) in zip(golden, CASES):
        u1, u2, groups = make_case(seed, **kw)
        assert c["name"] == name and c["umi1"] == u1 and c["umi2"] == u2 and c["groups"] == groups


def test_port_matches_golden(golden):
    from oracle import umi as U
    for c in golden:
        u1, u2, groups, t1, t2 = c["umi1"], c["umi2"], c["groups"], c["threshold1"], c["threshold2"]
        assert attempt(lambda: [[x + 1 for x in l] for l in U.port_levdist(u1, t1)]) == c["levdist"], c["name"]
        assert attempt(U.port_umi_group, u1, t1) == c["one"], c["name"]
        assert attempt(U.port_umi_group, u1, t1, None, None, groups) == c["one_grouped"], c["name"]
        assert attempt(U.port_umi_group, u1, t1, u2, t2, groups) == c["two_grouped"], c["name"]


def test_reference_matches_golden(golden, umiref):
    for c in golden:
        u1, u2, groups, t1, t2 = c["umi1"], c["umi2"], c["groups"], c["threshold1"], c["threshold2"]
        assert attempt(umiref.levdist, u1, t1, True) == c["levdist"]
        assert attempt(umiref.umi_group, u1, t1, u2, t2, groups) == c["two_grouped"]


def test_known_answers(umiref):
    from oracle import umi as U
    # N is half a mismatch against anything, N itself included (src/sorted_trie.cpp:15-21)
    assert U.lev2("ACGT", "ACGT") == 0 and U.lev2("ACGT", "ACGA") == 2 and U.lev2("ACNT", "ACGT") == 1
    assert U.lev2("NNNN", "NNNN") == 4 and U.lev2("ACGT", "ACG") == 2 and U.lev2("", "AC") == 4
    # trie order: ACGA < ACGT(1) < ACGT(4); the solo read comes first (src/cluster_umis.cpp:20-45)
    for f in (umiref.umi_group, U.port_umi_group):
        assert f(["ACGT", "ACGA", "TTTT", "ACGT"], 1) == [[3], [2, 1, 4]]
    # a read masked so heavily that it is not within the limit of itself
    for f in (umiref.umi_group, U.port_umi_group):
        with pytest.raises(RuntimeError, match="zero length read group"):
            f(["NNNNNNNN", "ACGTACGT"], 1)
    with pytest.raises(RuntimeError, match="should have the same length"):
        U.port_umi_group(["A", "C"], 1, ["A"], 1)


def test_clustering_against_the_reference_tests_slow_version(umiref):
    """tests/testthat/test-umicluster.R:4-29 (REF) and :32-41 (MOCKUP), compared as sets like COMPARE (:43-49)."""
    from oracle import umi as U
    rng = np.random.default_rng(5)

    def slow(groups):
        groups = [list(g) for g in groups]
        out = []
        for _ in range(len(groups)):
            sizes = [len(g) for g in groups]
            if max(sizes) == 0:
                break
            chosen = max(i for i, s in enumerate(sizes) if s == max(sizes))      # last, if ties
            cur = groups[chosen]
            out.append(cur)
            for j in cur:
                groups[j] = []
            groups = [[x for x in g if x not in cur] for g in groups]
        return out

    for nn, dens in ((20, 0.05), (20, 0.1), (20, 0.2), (50, 0.2), (50, 0.4), (50, 0.1), (50, 0.0)):
        m = rng.random((nn, nn)) < dens / 2
        m = m | m.T
        np.fill_diagonal(m, True)
        links = [np.nonzero(m[:, j])[0].tolist() for j in range(nn)]
        ref = slow(links)
        port = U.port_cluster(links)
        real = [[x - 1 for x in c] for c in umiref.cluster([[x + 1 for x in l] for l in links])]
        key = lambda cl: sorted(tuple(sorted(c)) for c in cl)   # noqa: E731
        assert key(port) == key(ref) == key(real)
        assert port == real          # and the exact order, which only the C++ defines


def test_host_clustering_entry_matches_the_reference(umiref):
    """sarlacc_cluster_umis (the library's host clustering, no device) == the reference's cluster_umis_test on random
    symmetric link sets (tests/testthat/test-umicluster.R:32-41), on asymmetric ones, and on its two error conditions."""
    from sarlacc_b200 import native, SarlaccError
    rng = np.random.default_rng(17)
    for nn, dens, symmetric in ((20, 0.05, True), (50, 0.2, True), (50, 0.4, True), (200, 0.03, True), (300, 0.0, True), (60, 0.1, False)):
        m = rng.random((nn, nn)) < dens / 2
        if symmetric:
            m = m | m.T
        np.fill_diagonal(m, True)
        links = [(np.nonzero(m[:, j])[0] + 1).tolist() for j in range(nn)]
        for l in links:
            rng.shuffle(l)                      # list order decides the order inside a cluster
        got = [c.tolist() for c in native.cluster_umis(links)]
        assert got == umiref.cluster(links)
    assert native.cluster_umis([]) == []
    with pytest.raises(SarlaccError, match="zero length read group"):
        native.cluster_umis([[1, 2], []])
    with pytest.raises(SarlaccError, match="single-read groups should contain only the read itself"):
        native.cluster_umis([[1, 2], [1]])
    with pytest.raises(SarlaccError, match="out of range"):
        native.cluster_umis([[1, 3], [2]])
