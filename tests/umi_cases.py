"""Seeded UMI test inputs shared by the CPU oracle tests, the golden generator and the GPU parity tests.
SEQSIM follows tests/testthat/test-umicluster.R:94-104 (a random reference, 10 % substitutions per copy), extended with
indels, N masking and junk characters to reach the corners of src/sorted_trie.cpp."""
import numpy as np


def seqsim(rng, n, length, p_sub=0.1, p_n=0.0, p_indel=0.0, junk=False):
    ref = rng.choice(list("ACGT"), length)
    out = []
    for _ in range(n):
        t = ref.copy()
        ch = rng.random(length) < p_sub
        t[ch] = rng.choice(list("ACGT"), int(ch.sum()))
        t[rng.random(length) < p_n] = "N"
        t = list(t)
        if p_indel and rng.random() < p_indel and len(t) > 1:
            del t[int(rng.integers(0, len(t)))]
        if p_indel and rng.random() < p_indel:
            t.insert(int(rng.integers(0, len(t) + 1)), str(rng.choice(list("ACGT"))))
        if junk and rng.random() < 0.03:
            t[int(rng.integers(0, len(t)))] = str(rng.choice(list("RYacgt-")))
        out.append("".join(t))
    return out


def make_case(seed, ngroups=6, lo=5, hi=40, length=10, **kw):
    """Returns (umi1, umi2, groups) -- groups as a list of 1-based index lists, reads shuffled across groups."""
    rng = np.random.default_rng(seed)
    u1, u2, label = [], [], []
    for g in range(ngroups):
        m = int(rng.integers(lo, hi + 1))
        for _ in range(int(rng.integers(1, 4))):          # a few molecules per pre-group
            k = max(1, m // 3)
            u1 += seqsim(rng, k, length, **kw)
            u2 += seqsim(rng, k, max(4, length - 2), **kw)
            label += [g] * k
    perm = rng.permutation(len(u1))
    u1 = [u1[i] for i in perm]
    u2 = [u2[i] for i in perm]
    label = np.asarray(label)[perm]
    groups = [(np.nonzero(label == g)[0] + 1).tolist() for g in range(ngroups) if (label == g).any()]
    return u1, u2, groups


CASES = [
    # name, seed, kwargs, threshold1, threshold2
    ("plain_t1", 11, dict(), 1, 1),
    ("plain_t3", 12, dict(), 3, 2),
    ("indels", 13, dict(p_indel=0.3), 2, 2),
    ("masked", 14, dict(p_n=0.08, p_indel=0.1), 3, 3),
    ("heavy_mask", 15, dict(p_n=0.3), 2, 1),              # self-distance can exceed the limit: the reference raises
    ("junk", 16, dict(junk=True, p_indel=0.1), 2, 2),     # characters the trie never stores
    ("long", 17, dict(length=40, p_indel=0.2, p_sub=0.05), 4, 4),
    ("tiny_groups", 18, dict(ngroups=30, lo=1, hi=3), 2, 2),
]
