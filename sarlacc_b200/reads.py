"""ReadSet: the host-side stand-in for Biostrings' QualityScaledDNAStringSet on this path.

A CSR container (byte pools + offsets) with exactly the operations the reference's R drivers apply to
reads around the hot path: width, subseq, reverseComplement, subsetting, names
(R/adaptorAlign.R:86-95,104-110,160-174).  Everything is vectorised numpy; nothing here aligns.
"""

import numpy as np

# Biostrings::reverseComplement on DNA with IUPAC codes
_COMP = np.arange(256, dtype=np.uint8)
for _a, _b in ["AT", "CG", "MK", "RY", "VB", "HD", "WW", "SS", "NN"]:
    _COMP[ord(_a)] = ord(_b)
    _COMP[ord(_b)] = ord(_a)
    _COMP[ord(_a.lower())] = ord(_b.lower())
    _COMP[ord(_b.lower())] = ord(_a.lower())


def _csr(strings):
    bs = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in strings]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        np.cumsum([len(b) for b in bs], out=off[1:])
    pool = np.frombuffer(b"".join(bs), dtype=np.uint8).copy()
    return pool, off


class ReadSet:
    """n sequences with per-base qualities (ASCII-encoded, e.g. Phred+33) and optional names."""

    def __init__(self, seq_pool, seq_off, qual_pool=None, qual_off=None, names=None):
        self.seq_pool = np.ascontiguousarray(seq_pool, dtype=np.uint8)
        self.seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
        self.qual_pool = None if qual_pool is None else np.ascontiguousarray(qual_pool, dtype=np.uint8)
        self.qual_off = self.seq_off if qual_off is None else np.ascontiguousarray(qual_off, dtype=np.int64)
        self.names = None if names is None else list(names)

    # -- construction --------------------------------------------------------------------------
    @classmethod
    def from_strings(cls, seqs, quals=None, names=None):
        sp, so = _csr(seqs)
        if quals is None:
            return cls(sp, so, None, None, names)
        qp, qo = _csr(quals)
        if len(qo) != len(so):
            raise ValueError("sequence and quality vectors should have the same length")
        return cls(sp, so, qp, qo, names)

    @classmethod
    def empty(cls):
        return cls(np.zeros(0, np.uint8), np.zeros(1, np.int64), np.zeros(0, np.uint8), np.zeros(1, np.int64), [])

    # -- basic accessors -----------------------------------------------------------------------
    def __len__(self):
        return len(self.seq_off) - 1

    @property
    def has_quality(self):
        return self.qual_pool is not None

    def width(self):
        return np.diff(self.seq_off).astype(np.int64)

    def seq_strings(self):
        b = self.seq_pool.tobytes()
        o = self.seq_off
        return [b[o[i]:o[i + 1]].decode("latin-1") for i in range(len(self))]

    def qual_strings(self):
        b = self.qual_pool.tobytes()
        o = self.qual_off
        return [b[o[i]:o[i + 1]].decode("latin-1") for i in range(len(self))]

    # -- R-level operations --------------------------------------------------------------------
    def __getitem__(self, idx):
        """reads[idx] with an index array or boolean mask."""
        idx = np.asarray(idx)
        if idx.dtype == bool:
            idx = np.nonzero(idx)[0]
        w = self.width()[idx]
        starts = np.ones(len(idx), dtype=np.int64)
        out = self._gather(idx, starts, w)
        if self.names is not None:
            out.names = [self.names[i] for i in idx]
        return out

    def _gather(self, rows, starts, widths):
        """New ReadSet whose element k is subseq(self[rows[k]], start=starts[k], width=widths[k]) (1-based)."""
        rows = np.asarray(rows, dtype=np.int64)
        starts = np.asarray(starts, dtype=np.int64)
        widths = np.asarray(widths, dtype=np.int64)
        off = np.zeros(len(rows) + 1, dtype=np.int64)
        np.cumsum(widths, out=off[1:])
        total = int(off[-1])
        src0 = self.seq_off[rows] + starts - 1
        idx = np.repeat(src0 - off[:-1], widths) + np.arange(total, dtype=np.int64)
        sp = self.seq_pool[idx]
        qp = None
        if self.has_quality:
            if self.qual_off is self.seq_off:
                qp = self.qual_pool[idx]
            else:
                qsrc0 = self.qual_off[rows] + starts - 1
                qidx = np.repeat(qsrc0 - off[:-1], widths) + np.arange(total, dtype=np.int64)
                qp = self.qual_pool[qidx]
        return ReadSet(sp, off, qp, off if qp is not None else None, None)

    def subseq(self, start=None, end=None, width=None):
        """XVector::subseq with vector arguments (two of start/end/width)."""
        w = self.width()
        n = len(self)
        if start is not None and width is not None:
            start = np.broadcast_to(np.asarray(start, dtype=np.int64), (n,))
            width = np.broadcast_to(np.asarray(width, dtype=np.int64), (n,))
        elif end is not None and width is not None:
            end = np.broadcast_to(np.asarray(end, dtype=np.int64), (n,))
            width = np.broadcast_to(np.asarray(width, dtype=np.int64), (n,))
            start = end - width + 1
        elif start is not None and end is not None:
            start = np.broadcast_to(np.asarray(start, dtype=np.int64), (n,))
            end = np.broadcast_to(np.asarray(end, dtype=np.int64), (n,))
            width = end - start + 1
        else:
            raise ValueError("two of start, end, width are required")
        if np.any(start < 1) or np.any(width < 0) or np.any(start + width - 1 > w):
            raise ValueError("subseq: requested range is out of bounds")
        out = self._gather(np.arange(n, dtype=np.int64), start, width)
        out.names = self.names
        return out

    def reverse_complement(self):
        """Biostrings::reverseComplement on a QualityScaledDNAStringSet: bases complemented and reversed,
        qualities reversed with them."""
        n = len(self)
        w = self.width()
        total = int(self.seq_off[-1] - self.seq_off[0])
        # position p of element k maps to off[k] + off[k+1] - 1 - p
        base = np.repeat(self.seq_off[:-1] + self.seq_off[1:] - 1, w)
        idx = base - (np.arange(total, dtype=np.int64) + self.seq_off[0])
        sp = _COMP[self.seq_pool[idx]]
        off = self.seq_off - self.seq_off[0]
        qp = None
        qo = None
        if self.has_quality:
            if self.qual_off is self.seq_off or np.array_equal(self.qual_off, self.seq_off):
                qp = self.qual_pool[idx]
            else:
                qw = np.diff(self.qual_off)
                qbase = np.repeat(self.qual_off[:-1] + self.qual_off[1:] - 1, qw)
                qidx = qbase - (np.arange(int(qw.sum()), dtype=np.int64) + self.qual_off[0])
                qp = self.qual_pool[qidx]
            qo = off
        return ReadSet(sp, off, qp, qo, self.names)

    @staticmethod
    def concat(sets):
        sets = list(sets)
        if not sets:
            return ReadSet.empty()
        sp = np.concatenate([s.seq_pool[s.seq_off[0]:s.seq_off[-1]] for s in sets])
        so = np.zeros(sum(len(s) for s in sets) + 1, dtype=np.int64)
        np.cumsum(np.concatenate([s.width() for s in sets]), out=so[1:])
        qp = None
        if all(s.has_quality for s in sets):
            qp = np.concatenate([s.qual_pool[s.qual_off[0]:s.qual_off[-1]] for s in sets])
        names = None
        if all(s.names is not None for s in sets):
            names = [x for s in sets for x in s.names]
        return ReadSet(sp, so, qp, so if qp is not None else None, names)


def _read_fastq_native(path, number):
    """Plain-text FASTQ through the library's buffered reader (sarlacc_fastq_*)."""
    import ctypes as C
    from . import _lib
    h = _lib.lib.sarlacc_fastq_open(str(path).encode())
    if not h:
        raise _lib.SarlaccError(_lib.last_error())
    try:
        ptrs = [C.c_void_p() for _ in range(6)]
        while True:
            n = _lib.lib.sarlacc_fastq_next(h, C.c_int64(int(number) if number else 1 << 62), *[C.byref(p) for p in ptrs])
            if n < 0:
                raise _lib.SarlaccError(_lib.last_error())
            if n == 0:
                return

            def arr(ptr, count, dtype):
                ctype = C.c_uint8 if dtype == np.uint8 else C.c_int64
                return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(max(count, 1),))[:count].copy()

            so = arr(ptrs[1], n + 1, np.int64)
            qo = arr(ptrs[3], n + 1, np.int64)
            no = arr(ptrs[5], n + 1, np.int64)
            sp = arr(ptrs[0], int(so[-1]), np.uint8)
            qp = arr(ptrs[2], int(qo[-1]), np.uint8)
            nb = arr(ptrs[4], int(no[-1]), np.uint8).tobytes()
            names = [nb[no[i]:no[i + 1]].decode("latin-1") for i in range(n)]
            yield ReadSet(sp, so, qp, qo, names)
    finally:
        _lib.lib.sarlacc_fastq_close(h)


def read_fastq_condensed(path, keep, number=None, nthreads=0):
    """Parallel FASTQ ingest for the adaptor path (sarlacc_fastq_next_condensed): yields (ReadSet, widths) where every
    read is cut down to its first and last `keep` bases (whole if no longer than 2*keep) and `widths` holds the true
    read lengths.  For tolerance <= keep the condensed reads have exactly the windows .get_front_and_back
    (R/adaptorAlign.R:86-95) cuts from the full reads."""
    import ctypes as C
    from . import _lib
    h = _lib.lib.sarlacc_fastq_open(str(path).encode())
    if not h:
        raise _lib.SarlaccError(_lib.last_error())
    try:
        ptrs = [C.c_void_p() for _ in range(7)]
        while True:
            n = _lib.lib.sarlacc_fastq_next_condensed(h, C.c_int64(int(number) if number else 1 << 62), C.c_int(int(keep)), C.c_int(int(nthreads)),
                                                      *[C.byref(p) for p in ptrs])
            if n < 0:
                raise _lib.SarlaccError(_lib.last_error())
            if n == 0:
                return

            def arr(ptr, count, ctype):
                return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(max(count, 1),))[:count].copy()

            so = arr(ptrs[1], n + 1, C.c_int64)
            qo = arr(ptrs[3], n + 1, C.c_int64)
            no = arr(ptrs[5], n + 1, C.c_int64)
            sp = arr(ptrs[0], int(so[-1]), C.c_uint8)
            qp = arr(ptrs[2], int(qo[-1]), C.c_uint8)
            nb = arr(ptrs[4], int(no[-1]), C.c_uint8).tobytes()
            widths = arr(ptrs[6], n, C.c_int32)
            names = [nb[no[i]:no[i + 1]].decode("latin-1") for i in range(n)]
            yield ReadSet(sp, so, qp, qo, names), widths
    finally:
        _lib.lib.sarlacc_fastq_close(h)


def read_fastq(path, number=None, skip=0):
    """FASTQ reader standing in for ShortRead::FastqStreamer + .FASTQ2QSDS (R/adaptorAlign.R:26,36,104-110).
    Yields ReadSets of at most `number` reads through the library's reader; gzip-compressed files are inflated by the
    library as they are read (zlib), like ShortRead does."""
    yield from _read_fastq_native(path, number)


def write_fastq(path, reads, append=False):
    s = reads.seq_strings()
    q = reads.qual_strings()
    names = reads.names if reads.names is not None else ["READ_%d" % (i + 1) for i in range(len(reads))]
    with open(path, "ab" if append else "wb") as fh:
        for i in range(len(reads)):
            fh.write(("@%s\n%s\n+\n%s\n" % (names[i], s[i], q[i])).encode("latin-1"))
