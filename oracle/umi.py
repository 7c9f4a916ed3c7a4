"""TEST INFRASTRUCTURE ONLY -- not part of the product.

CPU oracles for the UMI-grouping row:

* `UmiRef`  -- the reference's own umi_group / fast_levdist_test / cluster_umis_test, compiled verbatim into
               oracle/_ref/libsarlacc_umi_ref.so (oracle/Makefile `umiref`, oracle/umi_ref_driver.cpp).
* `port_*`  -- a plain-Python restatement for small cases (brute-force distance instead of the trie), every function
               citing the reference lines it follows.  tests/test_umi_oracle.py checks port == ref.
Indices are 1-based wherever R's are.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libsarlacc_umi_ref.so")


def _pool(strings):
    b = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in strings]
    off = np.zeros(len(b) + 1, np.int64)
    if b:
        off[1:] = np.cumsum([len(x) for x in b])
    pool = np.frombuffer(b"".join(b) + b"\0", np.uint8).copy()
    return pool, off


def _csr(lists):
    off = np.zeros(len(lists) + 1, np.int64)
    if len(lists):
        off[1:] = np.cumsum([len(x) for x in lists])
    vals = np.asarray([v for x in lists for v in x] + [0], np.int32)
    return off, vals


class UmiRef:
    @staticmethod
    def available():
        return os.path.exists(REF_LIB)

    def __init__(self):
        self.lib = C.CDLL(REF_LIB)

    def _fetch(self, rc, nl, nv, err):
        if rc != 0:
            raise RuntimeError(err.value.decode("latin-1"))
        off = np.zeros(nl.value + 1, np.int64)
        vals = np.zeros(max(nv.value, 1), np.int32)
        self.lib.ref_fetch(off.ctypes.data_as(C.c_void_p), vals.ctypes.data_as(C.c_void_p))
        return [vals[off[i]:off[i + 1]].tolist() for i in range(nl.value)]

    def umi_group(self, umi1, threshold1=3, umi2=None, threshold2=None, groups=None):
        """R/umiGroup.R:2-23 without the quality masking; groups = list of 1-based index lists (default: one group)."""
        n = len(umi1)
        if threshold2 is None:
            threshold2 = threshold1
        if groups is None:
            groups = [list(range(1, n + 1))]
        p1, o1 = _pool(umi1)
        if umi2 is not None:
            p2, o2 = _pool(umi2)
            a2, b2 = p2.ctypes.data_as(C.c_void_p), o2.ctypes.data_as(C.c_void_p)
        else:
            a2 = b2 = None
        go, gm = _csr(groups)
        nl, nv = C.c_int64(0), C.c_int64(0)
        err = C.create_string_buffer(512)
        rc = self.lib.ref_umi_group(p1.ctypes.data_as(C.c_void_p), o1.ctypes.data_as(C.c_void_p), C.c_int64(n), C.c_int(int(threshold1)),
                                    a2, b2, C.c_int(int(threshold2)), go.ctypes.data_as(C.c_void_p), gm.ctypes.data_as(C.c_void_p),
                                    C.c_int64(len(groups)), C.byref(nl), C.byref(nv), err, C.c_int(512))
        return self._fetch(rc, nl, nv, err)

    def levdist(self, seqs, limit, sorted_order=True):
        p, o = _pool(seqs)
        nl, nv = C.c_int64(0), C.c_int64(0)
        err = C.create_string_buffer(512)
        rc = self.lib.ref_levdist(p.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), C.c_int64(len(seqs)), C.c_int(int(limit)),
                                  C.c_int(1 if sorted_order else 0), C.byref(nl), C.byref(nv), err, C.c_int(512))
        return self._fetch(rc, nl, nv, err)

    def cluster(self, links):
        o, v = _csr(links)
        nl, nv = C.c_int64(0), C.c_int64(0)
        err = C.create_string_buffer(512)
        rc = self.lib.ref_cluster_umis(o.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), C.c_int64(len(links)),
                                       C.byref(nl), C.byref(nv), err, C.c_int(512))
        return self._fetch(rc, nl, nv, err)


# ---------------------------------------------------------------------------------------------- restatement
TRIE_ALPHABET = "ACGTN"      # src/sorted_trie.cpp:10 (children are visited in this order, :178-183)


def edit_score(a, b):
    """get_edit_score, src/sorted_trie.cpp:15-21: N against anything (N included) is half a mismatch."""
    if a == "N" or b == "N":
        return 1
    return 0 if a == b else 2


def lev2(s, t):
    """Twice the masked Levenshtein distance: the DP of src/sorted_trie.cpp:134-138 (indel = mismatch = 2)."""
    prev = [2 * i for i in range(len(s) + 1)]                       # :196-202
    for d, tb in enumerate(t, 1):
        cur = [prev[0] + 2] + [0] * len(s)                          # :111
        for i in range(1, len(s) + 1):
            cur[i] = min(prev[i] + 2, cur[i - 1] + 2, prev[i - 1] + edit_score(s[i - 1], tb))
        prev = cur
    return prev[len(s)]


def trie_key(s):
    return tuple(TRIE_ALPHABET.index(c) for c in s)


def port_levdist(seqs, limit):
    """What sorted_trie::find returns for every sequence (0-based lists): the stored sequences within `limit`, in the
    order the trie walk meets them -- a node's own indices (insertion order) before its children, children in ACGTN
    order (src/sorted_trie.cpp:152-185), i.e. sorted by (sequence under ACGTN, index).  Sequences with any other
    character are never stored (the switch of :53-69 has no default) but can still be queried."""
    stored = [i for i, s in enumerate(seqs) if all(c in TRIE_ALPHABET for c in s)]
    stored.sort(key=lambda i: (trie_key(seqs[i]), i))
    return [[j for j in stored if lev2(seqs[i], seqs[j]) <= 2 * limit] for i in range(len(seqs))]


def port_cluster(storage):
    """cluster_umis, src/cluster_umis.cpp:7-112, 0-based.  Solo reads first (:20-45), then repeatedly the node with
    the most unclaimed neighbours, LAST index on ties (:62-69), absorbing its unclaimed neighbours in list order and
    decrementing the counts of their neighbours (:76-100)."""
    n = len(storage)
    remaining = [len(x) for x in storage]
    out = []
    in_play = []
    for a in range(n):
        if remaining[a] > 1:
            in_play.append(a)
        elif remaining[a] == 1:
            if storage[a][0] != a:
                raise RuntimeError("single-read groups should contain only the read itself")
            out.append([a])
        else:
            raise RuntimeError("zero length read group")
    in_play = set(in_play)
    while True:
        live = [a for a in in_play if remaining[a] > 0]
        if not live:
            break
        top = max(live, key=lambda a: (remaining[a], a))
        in_play.discard(top)
        cluster = []
        for nb in storage[top]:
            if remaining[nb] == 0:
                continue
            cluster.append(nb)
            remaining[nb] = 0
            for nx in storage[nb]:
                if remaining[nx] > 0:
                    remaining[nx] -= 1
        out.append(cluster)
    return out


def port_umi_group(umi1, threshold1=3, umi2=None, threshold2=None, groups=None):
    """umi_group, src/umi_group.cpp:14-117 (+ unlist, R/umiGroup.R:22).  1-based in and out."""
    n = len(umi1)
    if threshold2 is None:
        threshold2 = threshold1
    if umi2 is not None and len(umi2) != n:
        raise RuntimeError("'umi1' and 'umi2' should have the same length")
    if groups is None:
        groups = [list(range(1, n + 1))]
    out = []
    for g in groups:
        if len(g) == 1:                                             # :39-42
            out.append(list(g))
            continue
        s1 = [umi1[i - 1] for i in g]
        m1 = port_levdist(s1, threshold1)
        if umi2 is None:
            storage = m1
        else:                                                       # :65-101: trie2 order, kept if also a UMI1 match
            s2 = [umi2[i - 1] for i in g]
            m2 = port_levdist(s2, threshold2)
            storage = [[x for x in m2[k] if x in set(m1[k])] for k in range(len(g))]
        for cl in port_cluster(storage):
            out.append([g[x] for x in cl])                          # :105-109
    return out
