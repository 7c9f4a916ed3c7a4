"""Python face of the four `.Call` entry points (and the fused/resident extensions), one function per
reference symbol, same argument order and meaning:

    cxx_adaptor_align            src/adaptor_align.cpp:11-77     -> adaptor_align
    cxx_adaptor_align_score_only src/adaptor_align.cpp:79-110    -> adaptor_align_score_only
    cxx_barcode_align            src/barcode_align.cpp:10-44     -> barcode_align
    cxx_general_align            src/general_align.cpp:10-62     -> general_align

The scalar/shape checks the reference performs on its SEXP arguments before the read loop
(src/utils.cpp:5-31, src/adaptor_align.cpp:23-31) are made here with the same messages, because in the
R integration they live in the glue (sarlacc_b200/csrc/r_glue.cpp), not behind the C ABI.
"""
import ctypes as C
import numbers

import numpy as np

from . import _lib
from ._lib import SarlaccError, SEQ_ASCII, SEQ_BIOSTRINGS  # noqa: F401
from .reads import ReadSet


def phred_encoding(n=94, offset=33):
    """What .create_encoding_vector (R/qualityMask.R:19-27) returns for PhredQuality input: names
    '!'..'~' and error probabilities 10^(-q/10)."""
    names = [chr(offset + i) for i in range(n)]
    err = np.array([10.0 ** (-q / 10.0) for q in range(n)], dtype=np.float64)
    return names, err


def _numeric_scalar(x, what):
    # check_numeric_scalar, src/utils.cpp:18-20
    if isinstance(x, numbers.Real):
        return float(x)
    a = np.asarray(x)
    if a.size != 1:
        raise SarlaccError("%s should be a numeric scalar" % what)
    return float(a.reshape(-1)[0])


def _string(x, what):
    # check_string, src/utils.cpp:26-31
    if isinstance(x, (str, bytes)):
        return x if isinstance(x, str) else x.decode("latin-1")
    x = list(x)
    if len(x) != 1:
        raise SarlaccError("%s should be a string" % what)
    return _string(x[0], what)


def _reads_arg(reads, views, seq_encoding):
    if not isinstance(reads, ReadSet):
        seqs, quals = reads
        sp, so = (seqs if isinstance(seqs, tuple) else _pool(seqs))
        qp, qo = (quals if isinstance(quals, tuple) else _pool(quals))
        if len(so) != len(qo):
            raise SarlaccError("sequence and quality vectors should have the same length")
        return _lib.ReadsArg(sp, so, qp, qo, seq_encoding, views)
    if not reads.has_quality:
        raise SarlaccError("sequence and quality vectors should have the same length")
    return _lib.ReadsArg(reads.seq_pool, reads.seq_off, reads.qual_pool, reads.qual_off, seq_encoding, views)


def _pool(strings):
    rs = ReadSet.from_strings(strings)
    return rs.seq_pool, rs.seq_off


def _encoding_arg(encoding):
    names, err = encoding
    return _lib.EncodingArg(names, err)


def adaptor_align(reads, encoding, gapopen, gapext, adaptor, sec_starts=(), sec_ends=(), views=False, seq_encoding=SEQ_ASCII):
    """Returns [score, start, end, [sec_start...], [sec_width...]] like the R list of src/adaptor_align.cpp:71-74.
    sec_starts are 0-based and sec_ends 1-based (R/adaptorAlign.R:158)."""
    adaptor = _string(adaptor, "adaptor sequence")
    go = _numeric_scalar(gapopen, "gap opening penalty")
    ge = _numeric_scalar(gapext, "gap extension penalty")
    ra = _reads_arg(reads, views, seq_encoding)
    ss = np.ascontiguousarray(sec_starts, dtype=np.int32).reshape(-1)
    se = np.ascontiguousarray(sec_ends, dtype=np.int32).reshape(-1)
    if len(ss) != len(se):
        raise SarlaccError("section starts and ends should have the same length")
    ea = _encoding_arg(encoding)
    n, nsec = ra.n, len(ss)
    score = np.zeros(n, np.float64)
    start = np.zeros(n, np.int32)
    end = np.zeros(n, np.int32)
    sst = np.zeros((max(nsec, 1), max(n, 1)), np.int32)
    swd = np.zeros((max(nsec, 1), max(n, 1)), np.int32)
    _lib.check(_lib.lib.sarlacc_adaptor_align(
        ra.ref(), ea.ref(), C.c_double(go), C.c_double(ge), adaptor.encode("latin-1"),
        C.c_int(nsec), _lib._ptr(ss), _lib._ptr(se),
        _lib._ptr(score), _lib._ptr(start), _lib._ptr(end), _lib._ptr(sst), _lib._ptr(swd)))
    return [score, start, end, [sst[i, :n].copy() for i in range(nsec)], [swd[i, :n].copy() for i in range(nsec)]]


def _score_call(fn, reads, encoding, gapopen, gapext, reference, what, views, seq_encoding):
    reference = _string(reference, what)
    go = _numeric_scalar(gapopen, "gap opening penalty")
    ge = _numeric_scalar(gapext, "gap extension penalty")
    ra = _reads_arg(reads, views, seq_encoding)
    ea = _encoding_arg(encoding)
    score = np.zeros(ra.n, np.float64)
    _lib.check(fn(ra.ref(), ea.ref(), C.c_double(go), C.c_double(ge), reference.encode("latin-1"), _lib._ptr(score)))
    return score


def adaptor_align_score_only(reads, encoding, gapopen, gapext, adaptor, views=False, seq_encoding=SEQ_ASCII):
    return _score_call(_lib.lib.sarlacc_adaptor_align_score_only, reads, encoding, gapopen, gapext, adaptor,
                       "adaptor sequence", views, seq_encoding)


def barcode_align(reads, encoding, gapopen, gapext, reference, views=False, seq_encoding=SEQ_ASCII):
    return _score_call(_lib.lib.sarlacc_barcode_align, reads, encoding, gapopen, gapext, reference,
                       "barcode sequence", views, seq_encoding)


def general_align(reads, encoding, gapopen, gapext, reference, edit_only=False, views=False, seq_encoding=SEQ_ASCII):
    """Returns [score, edit distance, reference strings, query strings] (src/general_align.cpp:60)."""
    reference = _string(reference, "reference sequence")
    go = _numeric_scalar(gapopen, "gap opening penalty")
    ge = _numeric_scalar(gapext, "gap extension penalty")
    ra = _reads_arg(reads, views, seq_encoding)
    ea = _encoding_arg(encoding)
    n = ra.n
    score = np.zeros(n, np.float64)
    edit = np.zeros(n, np.int32)
    maxlen = int(np.max(np.diff(ra.seq_off))) if n else 0
    stride = maxlen + len(reference) + 2
    ref_aln = np.zeros((max(n, 1), stride), np.uint8)
    q_aln = np.zeros((max(n, 1), stride), np.uint8)
    _lib.check(_lib.lib.sarlacc_general_align(
        ra.ref(), ea.ref(), C.c_double(go), C.c_double(ge), reference.encode("latin-1"), C.c_int(1 if edit_only else 0),
        _lib._ptr(score), _lib._ptr(edit), _lib._ptr(ref_aln), _lib._ptr(q_aln), C.c_int64(stride)))
    if edit_only:
        return [score, edit, [], []]
    rs = [bytes(ref_aln[i]).split(b"\0", 1)[0].decode("latin-1") for i in range(n)]
    qs = [bytes(q_aln[i]).split(b"\0", 1)[0].decode("latin-1") for i in range(n)]
    return [score, edit, rs, qs]


def barcode_align_multi(reads, encoding, gapopen, gapext, barcodes, all_scores=False, views=False, seq_encoding=SEQ_ASCII):
    """All barcodes in one pass (R/barcodeAlign.R:20-35 fused).  Returns best_id (1-based, 0 = NA), best, next_best
    and, if requested, the [nbarcodes][n] score matrix."""
    go = _numeric_scalar(gapopen, "gap opening penalty")
    ge = _numeric_scalar(gapext, "gap extension penalty")
    ra = _reads_arg(reads, views, seq_encoding)
    ea = _encoding_arg(encoding)
    bs = [_string(b, "barcode sequence").encode("latin-1") for b in barcodes]
    arr = (C.c_char_p * max(len(bs), 1))(*bs)
    n = ra.n
    bid = np.zeros(n, np.int32)
    best = np.zeros(n, np.float64)
    nxt = np.zeros(n, np.float64)
    mat = np.zeros((max(len(bs), 1), max(n, 1)), np.float64) if all_scores else None
    _lib.check(_lib.lib.sarlacc_barcode_align_multi(
        ra.ref(), ea.ref(), C.c_double(go), C.c_double(ge), arr, C.c_int(len(bs)),
        _lib._ptr(bid), _lib._ptr(best), _lib._ptr(nxt), _lib._ptr(mat)))
    if all_scores:
        return bid, best, nxt, mat[:len(bs), :n]
    return bid, best, nxt


def adaptor_align_windows(front, back, encoding, gapopen, gapext, adaptor1, adaptor2, sec1=((), ()), sec2=((), ()),
                          read_width=None, views=False, seq_encoding=SEQ_ASCII, reuse=None):
    """Fused .align_AA_internal (+ adaptor2 flip when read_width is given), see sarlacc_adaptor_align_windows.
    sec1/sec2 = (0-based starts, 1-based ends).  Returns reversed (bool) and two result lists shaped like
    adaptor_align's: [score, start, end, [sec_start...], [sec_width...]].  reuse: a dict the caller keeps between calls
    of the same shape -- the output arrays live in it and are overwritten by the next call (no fresh 49 MB of
    first-touched pages per million reads)."""
    a1 = _string(adaptor1, "adaptor sequence")
    a2 = _string(adaptor2, "adaptor sequence")
    go = _numeric_scalar(gapopen, "gap opening penalty")
    ge = _numeric_scalar(gapext, "gap extension penalty")
    rf = _reads_arg(front, views, seq_encoding)
    rb = _reads_arg(back, views, seq_encoding)
    if rf.n != rb.n:
        raise SarlaccError("front and back windows should have the same length")
    ea = _encoding_arg(encoding)
    n = rf.n
    secs = []
    for st, en in (sec1, sec2):
        ss = np.ascontiguousarray(st, dtype=np.int32).reshape(-1)
        se = np.ascontiguousarray(en, dtype=np.int32).reshape(-1)
        if len(ss) != len(se):
            raise SarlaccError("section starts and ends should have the same length")
        secs.append((ss, se))
    width = None if read_width is None else np.ascontiguousarray(read_width, dtype=np.int32)
    key = (n, len(secs[0][0]), len(secs[1][0]))
    if reuse is not None and reuse.get("key") == key:
        rev, outs = reuse["rev"], reuse["outs"]
    else:
        rev = np.empty(max(n, 1), np.uint8)      # every element is written by the call (n > 0) -- no zero fill, no copies below
        outs = []
        for ss, se in secs:
            outs.append([np.empty(n, np.float64), np.empty(n, np.int32), np.empty(n, np.int32),
                         np.empty((max(len(ss), 1), max(n, 1)), np.int32), np.empty((max(len(ss), 1), max(n, 1)), np.int32)])
        if reuse is not None:
            reuse.update(key=key, rev=rev, outs=outs)
    _lib.check(_lib.lib.sarlacc_adaptor_align_windows(
        rf.ref(), rb.ref(), ea.ref(), C.c_double(go), C.c_double(ge), a1.encode("latin-1"), a2.encode("latin-1"),
        C.c_int(len(secs[0][0])), _lib._ptr(secs[0][0]), _lib._ptr(secs[0][1]),
        C.c_int(len(secs[1][0])), _lib._ptr(secs[1][0]), _lib._ptr(secs[1][1]),
        _lib._ptr(width), _lib._ptr(rev),
        *[_lib._ptr(x) for x in outs[0]], *[_lib._ptr(x) for x in outs[1]]))
    res = []
    for k, (ss, se) in enumerate(secs):
        o = outs[k]
        res.append([o[0], o[1], o[2], [o[3][i, :n] for i in range(len(ss))], [o[4][i, :n] for i in range(len(ss))]])
    return rev[:n].view(np.bool_), res[0], res[1]


def adaptor_align_reads(reads, tolerance, encoding, gapopen, gapext, adaptor1, adaptor2, sec1=((), ()), sec2=((), ()),
                        views=False, seq_encoding=SEQ_ASCII):
    """Whole reads in: windows cut / reverse-complemented by the packer (sarlacc_adaptor_align_reads).  Returns
    read_width, reversed and the two result lists (adaptor2 coordinates already flipped into read coordinates)."""
    a1 = _string(adaptor1, "adaptor sequence")
    a2 = _string(adaptor2, "adaptor sequence")
    go = _numeric_scalar(gapopen, "gap opening penalty")
    ge = _numeric_scalar(gapext, "gap extension penalty")
    rr = _reads_arg(reads, views, seq_encoding)
    ea = _encoding_arg(encoding)
    n = rr.n
    secs = []
    for st, en in (sec1, sec2):
        ss = np.ascontiguousarray(st, dtype=np.int32).reshape(-1)
        se = np.ascontiguousarray(en, dtype=np.int32).reshape(-1)
        if len(ss) != len(se):
            raise SarlaccError("section starts and ends should have the same length")
        secs.append((ss, se))
    width = np.zeros(max(n, 1), np.int32)
    rev = np.empty(max(n, 1), np.uint8)      # every element is written by the call (n > 0) -- no zero fill, no copies below
    outs = []
    for ss, se in secs:
        outs.append([np.empty(n, np.float64), np.empty(n, np.int32), np.empty(n, np.int32),
                     np.empty((max(len(ss), 1), max(n, 1)), np.int32), np.empty((max(len(ss), 1), max(n, 1)), np.int32)])
    _lib.check(_lib.lib.sarlacc_adaptor_align_reads(
        rr.ref(), C.c_int(int(tolerance)), ea.ref(), C.c_double(go), C.c_double(ge), a1.encode("latin-1"), a2.encode("latin-1"),
        C.c_int(len(secs[0][0])), _lib._ptr(secs[0][0]), _lib._ptr(secs[0][1]),
        C.c_int(len(secs[1][0])), _lib._ptr(secs[1][0]), _lib._ptr(secs[1][1]),
        _lib._ptr(width), _lib._ptr(rev),
        *[_lib._ptr(x) for x in outs[0]], *[_lib._ptr(x) for x in outs[1]]))
    res = []
    for k, (ss, se) in enumerate(secs):
        o = outs[k]
        res.append([o[0], o[1], o[2], [o[3][i, :n] for i in range(len(ss))], [o[4][i, :n] for i in range(len(ss))]])
    return width[:n], rev[:n].view(np.bool_), res[0], res[1]


def _string_pool(strings):
    if isinstance(strings, ReadSet):
        return np.ascontiguousarray(strings.seq_pool, np.uint8), np.ascontiguousarray(strings.seq_off, np.int64)
    if isinstance(strings, tuple):
        return np.ascontiguousarray(strings[0], np.uint8), np.ascontiguousarray(strings[1], np.int64)
    rs = ReadSet.from_strings(list(strings))
    return np.ascontiguousarray(rs.seq_pool, np.uint8), np.ascontiguousarray(rs.seq_off, np.int64)


def _umi_call(fn, umi1, threshold1, umi2, threshold2, groups, device):
    p1, o1 = _string_pool(umi1)
    n = len(o1) - 1
    if p1.size == 0:
        p1 = np.zeros(1, np.uint8)
    p2 = o2 = None
    if umi2 is not None:
        p2, o2 = _string_pool(umi2)
        if len(o2) - 1 != n:
            raise SarlaccError("'umi1' and 'umi2' should have the same length")     # src/umi_group.cpp:25-29
        if p2.size == 0:
            p2 = np.zeros(1, np.uint8)
    if groups is None:
        groups = [np.arange(1, n + 1, dtype=np.int32)]
    goff = np.zeros(len(groups) + 1, np.int64)
    if len(groups):
        goff[1:] = np.cumsum([len(g) for g in groups])
    gmem = np.ascontiguousarray(np.concatenate([np.asarray(g, np.int32).reshape(-1) for g in groups] + [np.zeros(1, np.int32)]))
    h = fn(_lib._ptr(p1), _lib._ptr(o1), C.c_int64(n), C.c_int(int(threshold1)), _lib._ptr(p2), _lib._ptr(o2), C.c_int(int(threshold2)),
           _lib._ptr(goff), _lib._ptr(gmem), C.c_int64(len(groups)), C.c_int(int(device)))
    if not h:
        raise SarlaccError(_lib.last_error())
    try:
        nl, nv = _lib.lib.sarlacc_lists_count(h), _lib.lib.sarlacc_lists_values(h)
        off = np.zeros(nl + 1, np.int64)
        vals = np.zeros(max(nv, 1), np.int32)
        _lib.check(_lib.lib.sarlacc_lists_fetch(h, _lib._ptr(off), _lib._ptr(vals)))
    finally:
        _lib.lib.sarlacc_lists_free(h)
    return [vals[off[i]:off[i + 1]] for i in range(nl)]


def _fetch_lists(h):
    if not h:
        raise SarlaccError(_lib.last_error())
    try:
        nl, nv = _lib.lib.sarlacc_lists_count(h), _lib.lib.sarlacc_lists_values(h)
        off = np.zeros(nl + 1, np.int64)
        vals = np.zeros(max(nv, 1), np.int32)
        _lib.check(_lib.lib.sarlacc_lists_fetch(h, _lib._ptr(off), _lib._ptr(vals)))
    finally:
        _lib.lib.sarlacc_lists_free(h)
    return [vals[off[i]:off[i + 1]] for i in range(nl)]


def cluster_umis(links):
    """.Call(cxx_cluster_umis_test, links) (src/cluster_umis_test.cpp:8-29): greedy clustering of 1-based neighbour lists
    on the host (no device needed).  Returns a list of int32 arrays of 1-based indices."""
    off = np.zeros(len(links) + 1, np.int64)
    if len(links):
        off[1:] = np.cumsum([len(x) for x in links])
    vals = np.ascontiguousarray(np.concatenate([np.asarray(x, np.int32).reshape(-1) for x in links] + [np.zeros(1, np.int32)]))
    return _fetch_lists(_lib.lib.sarlacc_cluster_umis(_lib._ptr(off), _lib._ptr(vals), C.c_int64(len(links))))


def umi_group(umi1, threshold1, umi2=None, threshold2=None, groups=None, device=0):
    """.Call(cxx_umi_group, UMI1, threshold1, UMI2, threshold2, by.group) + unlist(recursive=FALSE)
    (src/umi_group.cpp:14-117, R/umiGroup.R:21-22): list of int32 arrays of 1-based read indices, one per cluster.
    groups: list of 1-based index vectors (R's by.group); default one group of all reads."""
    return _umi_call(_lib.lib.sarlacc_umi_group, umi1, threshold1, umi2, threshold1 if threshold2 is None else threshold2, groups, device)


def umi_neighbors(umi1, threshold1, umi2=None, threshold2=None, groups=None, device=0):
    """The neighbour lists umi_group clusters (1-based, trie order); with the default single group this is
    .Call(cxx_fast_levdist_test, seqs, limit, TRUE) (src/sorted_trie.cpp:302-332)."""
    return _umi_call(_lib.lib.sarlacc_umi_neighbors, umi1, threshold1, umi2, threshold1 if threshold2 is None else threshold2, groups, device)


class _Fetched(list):
    """[score, start, end, [section starts], [section widths]] that remembers its buffers for reuse by Resident.fetch."""


class _FetchedScores(np.ndarray):
    """Score vector that remembers its buffers for reuse by Resident.fetch."""


class Resident:
    """Read windows packed once and kept in HBM (sarlacc_resident_*)."""

    MODE_SCORE_LOCAL, MODE_TRACE_LOCAL, MODE_SCORE_GLOBAL = 0, 1, 2

    def __init__(self, reads, encoding, device=0, views=False, seq_encoding=SEQ_ASCII, _handle=None):
        if _handle is not None:
            self.handle = _handle
            self.n = int(_lib.lib.sarlacc_resident_n(self.handle))
            self.nsec = 0
            return
        ra = _reads_arg(reads, views, seq_encoding)
        ea = _encoding_arg(encoding)
        self.handle = _lib.lib.sarlacc_resident_create(ra.ref(), ea.ref(), C.c_int(device))
        if not self.handle:
            raise SarlaccError(_lib.last_error())
        self.n = int(_lib.lib.sarlacc_resident_n(self.handle))
        self.nsec = 0

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib.sarlacc_resident_free(self.handle)
            self.handle = None

    __del__ = close

    def scrambled(self, seed=0, first_index=0, read_index=None, stream_id=0):
        """Device-side .scramble_input: a new Resident whose windows are keyed random permutations of this one's."""
        idx = None if read_index is None else np.ascontiguousarray(read_index, dtype=np.uint64)
        h = _lib.lib.sarlacc_resident_scrambled(self.handle, C.c_uint64(int(seed)), C.c_uint64(int(first_index)), _lib._ptr(idx), C.c_int(int(stream_id)))
        if not h:
            raise SarlaccError(_lib.last_error())
        return Resident(None, None, _handle=h)

    def rows(self):
        """Packed rows (uint16[n][stride]) and lengths, downloaded (tests)."""
        stride = C.c_int(0)
        _lib.check(_lib.lib.sarlacc_resident_rows(self.handle, None, None, C.byref(stride)))
        rows = np.zeros((max(self.n, 1), stride.value), np.uint16)
        lens = np.zeros(max(self.n, 1), np.int32)
        _lib.check(_lib.lib.sarlacc_resident_rows(self.handle, _lib._ptr(rows), _lib._ptr(lens), C.byref(stride)))
        return rows[:self.n], lens[:self.n]

    def cells(self, rlen):
        return int(_lib.lib.sarlacc_resident_cells(self.handle, C.c_int(rlen)))

    def nbytes(self):
        return int(_lib.lib.sarlacc_resident_bytes(self.handle))

    def align(self, mode, gapopen, gapext, reference, sec_starts=(), sec_ends=(), stream=None):
        ss = np.ascontiguousarray(sec_starts, dtype=np.int32).reshape(-1)
        se = np.ascontiguousarray(sec_ends, dtype=np.int32).reshape(-1)
        if len(ss) != len(se):
            raise SarlaccError("section starts and ends should have the same length")
        self.nsec = len(ss) if mode == self.MODE_TRACE_LOCAL else 0
        self.mode = mode
        _lib.check(_lib.lib.sarlacc_resident_align(
            self.handle, C.c_int(mode), C.c_double(float(gapopen)), C.c_double(float(gapext)),
            reference.encode("latin-1"), C.c_int(len(ss)), _lib._ptr(ss), _lib._ptr(se),
            C.c_void_p(stream) if stream else None))

    def fetch(self, stream=None, pinned=False, out=None):
        """Copies the last run's results to the host.  pinned=True puts them in page-locked memory (torch allocator):
        the device-to-host copy then runs at PCIe speed instead of through the driver's bounce buffers.  out = the
        object returned by an earlier fetch of the same shape: its buffers are reused (page-locking is slow)."""
        n, nsec = self.n, self.nsec
        trace = self.mode == self.MODE_TRACE_LOCAL

        def alloc(shape, dtype):
            if pinned:
                import torch
                return torch.empty(shape, dtype={np.float64: torch.float64, np.int32: torch.int32}[dtype], pin_memory=True).numpy()
            return np.empty(shape, dtype)

        bufs = getattr(out, "_buffers", None) if out is not None else None
        if bufs is not None and bufs[0].shape == (n,) and (bufs[3] is None) == (not trace) and (not trace or bufs[3].shape == (max(nsec, 1), max(n, 1))):
            score, start, end, sst, swd = bufs
        else:
            score = alloc(n, np.float64)
            start = alloc(n, np.int32) if trace else None
            end = alloc(n, np.int32) if trace else None
            sst = alloc((max(nsec, 1), max(n, 1)), np.int32) if trace else None
            swd = alloc((max(nsec, 1), max(n, 1)), np.int32) if trace else None
        _lib.check(_lib.lib.sarlacc_resident_fetch(
            self.handle, _lib._ptr(score), _lib._ptr(start), _lib._ptr(end), _lib._ptr(sst), _lib._ptr(swd),
            C.c_void_p(stream) if stream else None))
        res = _Fetched([score, start, end, [sst[i, :n] for i in range(nsec)], [swd[i, :n] for i in range(nsec)]]) if trace else score.view(_FetchedScores)
        res._buffers = (score, start, end, sst, swd)
        return res

    def set_timing(self, on=True):
        _lib.lib.sarlacc_resident_set_timing(self.handle, C.c_int(1 if on else 0))

    def forward_ms(self):
        return float(_lib.lib.sarlacc_resident_forward_ms(self.handle))

    def scores_device_ptr(self):
        return _lib.lib.sarlacc_resident_scores_device(self.handle)

    def last_kernel(self):
        return _lib.lib.sarlacc_resident_last_kernel(self.handle).decode()


def _out_ptr(x):
    """A destination the chunk calls accept: None, a numpy array (host), or an integer device address (e.g. torch's data_ptr())."""
    if x is None:
        return None
    if isinstance(x, (int, np.integer)):
        return C.c_void_p(int(x))
    return x.ctypes.data_as(C.c_void_p)


class Chunk:
    """Device-resident reads re-loaded in place (sarlacc_chunk_*): what one FastqStreamer yield() is to the R drivers.
    adaptor_align = .align_AA_internal (+ adaptor2 flip), scrambled_scores = .align_AT_internal.  Calls only enqueue work;
    outputs (numpy arrays -- ideally page-locked -- or integer device addresses) are complete after sync()."""

    def __init__(self, capacity, tolerance, encoding, device=0):
        ea = _encoding_arg(encoding)
        self.handle = _lib.lib.sarlacc_chunk_create(C.c_int(device), C.c_int64(int(capacity)), C.c_int(int(tolerance)), ea.ref())
        if not self.handle:
            raise SarlaccError(_lib.last_error())
        self.capacity, self.tolerance, self.device = int(capacity), int(tolerance), device
        self._keep = []

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib.sarlacc_chunk_free(self.handle)
            self.handle = None

    __del__ = close

    @property
    def n(self):
        return int(_lib.lib.sarlacc_chunk_n(self.handle))

    def load_windows(self, front, back, width=None, views=False, seq_encoding=SEQ_ASCII):
        rf, rb = _reads_arg(front, views, seq_encoding), _reads_arg(back, views, seq_encoding)
        w = None if width is None else np.ascontiguousarray(width, dtype=np.int32)
        _lib.check(_lib.lib.sarlacc_chunk_load_reads(self.handle, rf.ref(), rb.ref(), C.c_int(0), _lib._ptr(w)))

    def load_reads(self, reads, tolerance=None, views=False, seq_encoding=SEQ_ASCII):
        rr = _reads_arg(reads, views, seq_encoding)
        tol = self.tolerance if tolerance is None else int(tolerance)
        _lib.check(_lib.lib.sarlacc_chunk_load_reads(self.handle, rr.ref(), None, C.c_int(tol), None))

    def load_mock(self, n, adaptor1, adaptor2, seed=2000, first_index=0, insert_len=4908, barcodes=None,
                  sub_rate=0.05, indel_rate=0.01, max_insert=5):
        bs = [b.encode("latin-1") for b in (barcodes or [])]
        arr = (C.c_char_p * max(len(bs), 1))(*bs)
        _lib.check(_lib.lib.sarlacc_chunk_load_mock(
            self.handle, C.c_int64(int(n)), C.c_uint64(int(first_index)), C.c_uint64(int(seed)),
            adaptor1.encode("latin-1"), adaptor2.encode("latin-1"), C.c_int(int(insert_len)), arr if bs else None, C.c_int(len(bs)),
            C.c_double(sub_rate), C.c_double(indel_rate), C.c_int(int(max_insert))))

    def adaptor_align(self, gapopen, gapext, adaptor1, adaptor2, sec1=((), ()), sec2=((), ()), out=None, out_pitch=0):
        """out: dict with any of reversed, width, score1, start1, end1, sec_start1, sec_width1, score2, ... (numpy arrays or
        device addresses, already offset to this chunk's first read; section matrices have row pitch out_pitch).  Without
        `out`, fresh arrays are allocated, the call is synchronised and (width, reversed, result1, result2) returned."""
        secs = []
        for st, en in (sec1, sec2):
            ss = np.ascontiguousarray(st, dtype=np.int32).reshape(-1)
            se = np.ascontiguousarray(en, dtype=np.int32).reshape(-1)
            if len(ss) != len(se):
                raise SarlaccError("section starts and ends should have the same length")
            secs.append((ss, se))
        own = out is None
        n = self.n
        if own:
            out = {"reversed": np.empty(max(n, 1), np.uint8), "width": np.zeros(max(n, 1), np.int32)}
            for k, (ss, _) in enumerate(secs, 1):
                out["score%d" % k] = np.empty(n, np.float64)
                out["start%d" % k] = np.empty(n, np.int32)
                out["end%d" % k] = np.empty(n, np.int32)
                out["sec_start%d" % k] = np.empty((max(len(ss), 1), max(n, 1)), np.int32)
                out["sec_width%d" % k] = np.empty((max(len(ss), 1), max(n, 1)), np.int32)
            out_pitch = max(n, 1)
        self._keep = [secs, out]          # the library reads the section lists while it builds its plan only; outputs until sync
        g = out.get
        _lib.check(_lib.lib.sarlacc_chunk_adaptor_align(
            self.handle, C.c_double(float(gapopen)), C.c_double(float(gapext)), adaptor1.encode("latin-1"), adaptor2.encode("latin-1"),
            C.c_int(len(secs[0][0])), _lib._ptr(secs[0][0]), _lib._ptr(secs[0][1]),
            C.c_int(len(secs[1][0])), _lib._ptr(secs[1][0]), _lib._ptr(secs[1][1]),
            C.c_int64(int(out_pitch)), _out_ptr(g("width")), _out_ptr(g("reversed")),
            _out_ptr(g("score1")), _out_ptr(g("start1")), _out_ptr(g("end1")), _out_ptr(g("sec_start1")), _out_ptr(g("sec_width1")),
            _out_ptr(g("score2")), _out_ptr(g("start2")), _out_ptr(g("end2")), _out_ptr(g("sec_start2")), _out_ptr(g("sec_width2"))))
        if not own:
            return None
        self.sync()
        res = []
        for k, (ss, _) in enumerate(secs, 1):
            res.append([out["score%d" % k], out["start%d" % k], out["end%d" % k],
                        [out["sec_start%d" % k][i, :n] for i in range(len(ss))], [out["sec_width%d" % k][i, :n] for i in range(len(ss))]])
        return out["width"][:n], out["reversed"][:n].view(np.bool_), res[0], res[1]

    def scrambled_scores(self, gapopen, gapext, adaptor1, adaptor2, seed=0, first_index=0, read_index=None, scramble=True,
                         score1=None, score2=None, strand_score=None):
        """.align_AT_internal (scramble=True), the same on the windows scrambled by the previous call (scramble="reuse"), or
        .get_alignment_scores + .resolve_strand on the windows as loaded (scramble=False).  score1 / score2: the kept score
        per adaptor; strand_score: .resolve_strand()$scores.  Without destinations, fresh arrays are allocated, the call is
        synchronised and (score1, score2) returned."""
        own = score1 is None and score2 is None and strand_score is None
        n = self.n
        if own:
            score1, score2 = np.empty(n, np.float64), np.empty(n, np.float64)
        idx = None if read_index is None else np.ascontiguousarray(read_index, dtype=np.uint64)
        self._keep = [score1, score2, strand_score, idx]
        mode = 2 if scramble == "reuse" else (1 if scramble else 0)
        _lib.check(_lib.lib.sarlacc_chunk_scrambled_scores(
            self.handle, C.c_double(float(gapopen)), C.c_double(float(gapext)), adaptor1.encode("latin-1"), adaptor2.encode("latin-1"),
            C.c_uint64(int(seed)), C.c_uint64(int(first_index)), _lib._ptr(idx), C.c_int(mode),
            _out_ptr(score1), _out_ptr(score2), _out_ptr(strand_score)))
        if own:
            self.sync()
            return score1, score2
        return None

    def sync(self):
        _lib.check(_lib.lib.sarlacc_chunk_sync(self.handle))

    def join(self):
        """Compute stream waits for the traceback and copy streams (see sarlacc_chunk_join)."""
        _lib.check(_lib.lib.sarlacc_chunk_join(self.handle))

    def stream(self):
        """The compute stream's handle (for torch.cuda.ExternalStream)."""
        return int(_lib.lib.sarlacc_chunk_stream(self.handle) or 0)

    def rows(self, which=0):
        """(packed rows uint16[n][stride], window lengths, read widths, strand flips) of window set `which`
        (0 front, 1 back, 2 scrambled front, 3 scrambled back)."""
        stride = C.c_int(0)
        _lib.check(_lib.lib.sarlacc_chunk_rows(self.handle, C.c_int(which), None, None, C.byref(stride), None, None))
        n = self.n
        rows = np.zeros((max(n, 1), stride.value), np.uint16)
        lens = np.zeros(max(n, 1), np.int32)
        width = np.zeros(max(n, 1), np.int32)
        flipped = np.zeros(max(n, 1), np.uint8)
        _lib.check(_lib.lib.sarlacc_chunk_rows(self.handle, C.c_int(which), _lib._ptr(rows), _lib._ptr(lens), C.byref(stride),
                                              _lib._ptr(width), _lib._ptr(flipped)))
        return rows[:n], lens[:n], width[:n], flipped[:n].view(np.bool_)

    def set_timing(self, on=True):
        _lib.lib.sarlacc_chunk_set_timing(self.handle, C.c_int(1 if on else 0))

    def phase_ms(self):
        ms = np.zeros(4, np.float64)
        _lib.check(_lib.lib.sarlacc_chunk_phase_ms(self.handle, _lib._ptr(ms)))
        return dict(zip(("load", "adaptor_align", "scramble", "score_only"), ms.tolist()))

    def last_kernel(self, adaptor=0):
        return _lib.lib.sarlacc_chunk_last_kernel(self.handle, C.c_int(adaptor)).decode()


def last_pair_timing():
    """Host-side phases (ms) of the last fused both-ends host-buffer call: staging, enqueue, wait + copy-out, total."""
    ms = np.zeros(6, np.float64)
    _lib.lib.sarlacc_last_pair_timing(_lib._ptr(ms))
    out = dict(zip(("stage", "enqueue", "wait_copy_out", "total", "upload_sum", "kernels_copyback_sum"), ms.tolist()))
    out["upload_bytes"] = float(_lib.lib.sarlacc_last_pair_upload_bytes())      # window bytes actually put on the link
    return out


def compute_threshold(real, scrambled, error, device=0):
    """.compute_threshold (R/getAdaptorThresholds.R:94-103) on the device.  real / scrambled: numpy float64 arrays, or
    (device address, length) pairs for vectors already on `device`."""
    def arg(x):
        if isinstance(x, tuple):
            return C.c_void_p(int(x[0])), int(x[1]), None
        a = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        return _lib._ptr(a), len(a), a
    rp, rn, rk = arg(real)
    sp, sn, sk = arg(scrambled)
    out = np.zeros(1, np.float64)
    _lib.check(_lib.lib.sarlacc_compute_threshold(rp, C.c_int64(rn), sp, C.c_int64(sn), C.c_double(float(error)), C.c_int(device), _lib._ptr(out)))
    return float(out[0])


def tied_overlap(real, fake, device=0):
    """.tied_overlap (R/tuneAlignment.R:78-86) on the device; numpy arrays or (device address, length) pairs."""
    def arg(x):
        if isinstance(x, tuple):
            return C.c_void_p(int(x[0])), int(x[1]), None
        a = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        return _lib._ptr(a), len(a), a
    rp, rn, rk = arg(real)
    fp, fn, fk = arg(fake)
    out = np.zeros(1, np.float64)
    _lib.check(_lib.lib.sarlacc_tied_overlap(rp, C.c_int64(rn), fp, C.c_int64(fn), C.c_int(device), _lib._ptr(out)))
    return float(out[0])
