"""GPU parity for the UMI-grouping row: the device neighbour search + host clustering behind
sarlacc_umi_group / sarlacc_umi_neighbors against the reference's own code (oracle/_ref/libsarlacc_umi_ref.so when it
travelled, else the restatement) and the committed golden vectors.  Everything is integer work: exact equality."""
import json
import os

import numpy as np
import pytest

from umi_cases import CASES, make_case, seqsim

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def attempt(f, *a, **k):
    from sarlacc_b200 import SarlaccError
    try:
        return {"value": [list(map(int, x)) for x in f(*a, **k)]}
    except (RuntimeError, SarlaccError) as e:
        return {"error": str(e)}


@pytest.fixture(scope="module")
def checker():
    from oracle import umi as U
    if U.UmiRef.available():
        R = U.UmiRef()
        return R.umi_group, lambda s, t: R.levdist(s, t, True)
    return U.port_umi_group, lambda s, t: [[x + 1 for x in l] for l in U.port_levdist(s, t)]


def test_golden_vectors():
    from sarlacc_b200 import native
    with open(os.path.join(ROOT, "tests", "golden", "umi_vectors.json")) as fh:
        cases = json.load(fh)["cases"]
    for c in cases:
        u1, u2, groups, t1, t2 = c["umi1"], c["umi2"], c["groups"], c["threshold1"], c["threshold2"]
        assert attempt(native.umi_neighbors, u1, t1) == c["levdist"], c["name"]
        assert attempt(native.umi_group, u1, t1) == c["one"], c["name"]
        assert attempt(native.umi_group, u1, t1, groups=groups) == c["one_grouped"], c["name"]
        assert attempt(native.umi_group, u1, t1, u2, t2, groups=groups) == c["two_grouped"], c["name"]


def test_random_against_the_checker(checker):
    from sarlacc_b200 import native
    group_f, lev_f = checker
    for seed in range(100, 112):
        kw = [dict(), dict(p_indel=0.3), dict(p_n=0.05, p_indel=0.1), dict(junk=True)][seed % 4]
        u1, u2, groups = make_case(seed, ngroups=5, lo=3, hi=60, length=[8, 12, 16, 20][seed % 4], **kw)
        for t in (0, 1, 3):
            assert attempt(native.umi_neighbors, u1, t) == attempt(lev_f, u1, t)
            assert attempt(native.umi_group, u1, t, groups=groups) == attempt(group_f, u1, t, None, None, groups)
        assert attempt(native.umi_group, u1, 2, u2, 1, groups=groups) == attempt(group_f, u1, 2, u2, 1, groups)


def test_edges(checker):
    from sarlacc_b200 import native, SarlaccError
    group_f, lev_f = checker
    # empty UMIs, duplicates, one read, lengths up to the kernel's 64, a group of one and an empty group
    seqs = ["", "A", "", "ACGT", "ACGT", "N", "ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT"[:64], "ACG"]
    for t in (0, 1, 2, 5):
        assert attempt(native.umi_neighbors, seqs, t) == attempt(lev_f, seqs, t)
        assert attempt(native.umi_group, seqs, t) == attempt(group_f, seqs, t)
    groups = [[1, 2, 3], [4], [], [5, 6, 7, 8]]
    assert attempt(native.umi_group, seqs, 1, groups=groups) == attempt(group_f, seqs, 1, None, None, groups)
    assert attempt(native.umi_group, ["ACGT"], 1) == {"value": [[1]]}
    assert attempt(native.umi_group, [], 1, groups=[]) == {"value": []}
    with pytest.raises(SarlaccError, match="should have the same length"):
        native.umi_group(["A", "C"], 1, ["A"], 1)
    with pytest.raises(SarlaccError, match="zero length read group"):
        native.umi_group(["NNNNNNNN", "ACGTACGT"], 1)
    with pytest.raises(SarlaccError, match="outside the device kernel"):
        native.umi_group(["A" * 65, "C" * 65], 1)
    with pytest.raises(SarlaccError, match="out of range"):
        native.umi_group(["A", "C"], 1, groups=[[1, 3]])


def test_large_group_properties():
    """20 000 UMIs of 12 bases in one pre-group (4e8 pairs): every read lands in exactly one cluster, every cluster's
    members are within the threshold of its seed's list, and a strided sample of neighbour lists equals the brute-force
    restatement."""
    from sarlacc_b200 import native
    from oracle import umi as U
    rng = np.random.default_rng(9)
    mol = ["".join(rng.choice(list("ACGT"), 12)) for _ in range(2500)]
    seqs = []
    for m in mol:
        for _ in range(8):
            t = list(m)
            for k in np.nonzero(rng.random(12) < 0.03)[0]:
                t[k] = str(rng.choice(list("ACGT")))
            seqs.append("".join(t))
    perm = rng.permutation(len(seqs))
    seqs = [seqs[i] for i in perm]
    clusters = native.umi_group(seqs, 1)
    flat = np.concatenate(clusters)
    assert len(flat) == len(seqs) and len(np.unique(flat)) == len(seqs)
    nb = native.umi_neighbors(seqs, 1)
    assert len(nb) == len(seqs)
    order = sorted(range(len(seqs)), key=lambda i: (U.trie_key(seqs[i]), i))
    for i in range(0, len(seqs), 997):
        want = [j + 1 for j in order if abs(len(seqs[i]) - len(seqs[j])) <= 1 and U.lev2(seqs[i], seqs[j]) <= 2]
        assert nb[i].tolist() == want


def test_umiGroup_mirror(checker):
    """api.umiGroup: label vectors are split() like R (groups in sorted label order), max_err masks through qualityMask."""
    from sarlacc_b200 import api, ReadSet
    group_f, _ = checker
    rng = np.random.default_rng(3)
    u1, u2, groups = make_case(21, ngroups=4, lo=5, hi=30)
    labels = np.zeros(len(u1), np.int64)
    for g, members in enumerate(groups):
        labels[np.asarray(members) - 1] = 10 - g          # descending labels: split() reorders the groups
    by_label = [groups[g] for g in np.argsort([10 - g for g in range(len(groups))])]
    got = api.umiGroup(u1, threshold1=1, groups=labels)
    assert [x.tolist() for x in got] == group_f(u1, 1, None, None, by_label)
    got = api.umiGroup(u1, threshold1=2, UMI2=u2, threshold2=1, groups=groups)
    assert [x.tolist() for x in got] == group_f(u1, 2, u2, 1, groups)
    # masking: qualities below Q10 become N before grouping (R/umiGroup.R:8-11)
    quals = ["".join(chr(33 + int(q)) for q in rng.integers(2, 40, size=len(s))) for s in u1]
    rs = ReadSet.from_strings(u1, quals)
    masked = ["".join("N" if (ord(q) - 33) < 10 else b for b, q in zip(s, ql)) for s, ql in zip(u1, quals)]
    assert api.qualityMask(rs, 0.1).seq_strings() == masked
    assert attempt(api.umiGroup, rs, threshold1=3, max_err=0.1) == attempt(group_f, masked, 3)
