"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the two CPU oracles (see oracle_abi.h).

``Oracle("port")`` loads oracle/libsarlacc_oracle.so (plain-C restatement, sarlacc_oracle.c);
``Oracle("ref")`` loads oracle/_ref/libsarlacc_ref.so (the reference's own reference_align.cpp
compiled verbatim).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module; the product package never does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libsarlacc_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsarlacc_ref.so")


def phred_encoding(n=94, offset=33):
    """The vector .create_encoding_vector (R/qualityMask.R:19-27) yields for PhredQuality:
    names '!'.. , error probability 10^(-q/10)."""
    names = [chr(offset + i) for i in range(n)]
    err = np.array([10.0 ** (-q / 10.0) for q in range(n)], dtype=np.float64)
    return names, err


def to_csr(strings):
    """list of bytes/str -> (pool uint8 array, int64 offsets[n+1])."""
    bs = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in strings]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs])
    pool = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    if pool.size == 0:
        pool = np.zeros(1, np.uint8)  # keep a valid pointer
    return pool, off


class OracleError(RuntimeError):
    pass


class Oracle:
    def __init__(self, kind="port"):
        self.kind = kind
        path = PORT_SO if kind == "port" else REF_SO
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.p = "orc" if kind == "port" else "ref"

    @staticmethod
    def available(kind):
        return os.path.exists(PORT_SO if kind == "port" else REF_SO)

    # -- helpers -------------------------------------------------------------------------------
    @staticmethod
    def _enc(encoding):
        names, err = encoding
        err = np.ascontiguousarray(err, dtype=np.float64)
        if names is None:
            arr = None
        else:
            arr = (C.c_char_p * max(len(names), 1))(*[n.encode("latin-1") if isinstance(n, str) else n for n in names])
        return arr, err, len(err)

    @staticmethod
    def _csr(x):
        if isinstance(x, tuple):
            pool, off = x
            pool = np.ascontiguousarray(pool, dtype=np.uint8)
            if pool.size == 0:
                pool = np.zeros(1, np.uint8)
            return pool, np.ascontiguousarray(off, dtype=np.int64)
        return to_csr(x)

    def _call(self, fn, *args):
        err = C.create_string_buffer(256)
        rc = fn(*args, err, 256)
        if rc != 0:
            raise OracleError(err.value.decode())

    # -- entry points ----------------------------------------------------------------------------
    def adaptor_align(self, seqs, quals, encoding, go, ge, adaptor, sec_starts=(), sec_ends=(), nthreads=1):
        """cxx_adaptor_align (src/adaptor_align.cpp:11-77). sec_starts 0-based, sec_ends 1-based,
        as R passes them (R/adaptorAlign.R:158).  Returns score, start, end, sec_start[nsec][n], sec_width."""
        sp, so = self._csr(seqs)
        qp, qo = self._csr(quals)
        n = len(so) - 1
        if len(qo) - 1 != n:
            raise OracleError("sequence and quality vectors should have the same length")
        ss = np.ascontiguousarray(sec_starts, dtype=np.int32)
        se = np.ascontiguousarray(sec_ends, dtype=np.int32)
        if len(ss) != len(se):
            raise OracleError("section starts and ends should have the same length")
        nsec = len(ss)
        names, err, en = self._enc(encoding)
        score = np.zeros(n, np.float64)
        start = np.zeros(n, np.int32)
        end = np.zeros(n, np.int32)
        sst = np.zeros((nsec, n), np.int32)
        swd = np.zeros((nsec, n), np.int32)
        fn = getattr(self.lib, self.p + "_adaptor_align")
        fn.restype = C.c_int
        self._call(fn, C.c_int64(n), sp.ctypes.data_as(C.c_char_p), so.ctypes.data_as(C.c_void_p),
                   qp.ctypes.data_as(C.c_char_p), qo.ctypes.data_as(C.c_void_p),
                   C.c_int(en), names, err.ctypes.data_as(C.c_void_p), C.c_double(go), C.c_double(ge),
                   adaptor.encode("latin-1"), C.c_int(nsec), ss.ctypes.data_as(C.c_void_p), se.ctypes.data_as(C.c_void_p),
                   score.ctypes.data_as(C.c_void_p), start.ctypes.data_as(C.c_void_p), end.ctypes.data_as(C.c_void_p),
                   sst.ctypes.data_as(C.c_void_p), swd.ctypes.data_as(C.c_void_p), C.c_int(nthreads))
        return score, start, end, sst, swd

    def align_score_only(self, seqs, quals, encoding, go, ge, reference, local=True, nthreads=1):
        """cxx_adaptor_align_score_only (local=True, src/adaptor_align.cpp:79-110) or
        cxx_barcode_align (local=False, src/barcode_align.cpp:10-44)."""
        sp, so = self._csr(seqs)
        qp, qo = self._csr(quals)
        n = len(so) - 1
        if len(qo) - 1 != n:
            raise OracleError("sequence and quality vectors should have the same length")
        names, err, en = self._enc(encoding)
        score = np.zeros(n, np.float64)
        fn = getattr(self.lib, self.p + "_align_score_only")
        fn.restype = C.c_int
        self._call(fn, C.c_int64(n), sp.ctypes.data_as(C.c_char_p), so.ctypes.data_as(C.c_void_p),
                   qp.ctypes.data_as(C.c_char_p), qo.ctypes.data_as(C.c_void_p),
                   C.c_int(en), names, err.ctypes.data_as(C.c_void_p), C.c_double(go), C.c_double(ge),
                   reference.encode("latin-1"), C.c_int(1 if local else 0),
                   score.ctypes.data_as(C.c_void_p), C.c_int(nthreads))
        return score

    def general_align(self, seqs, quals, encoding, go, ge, reference, edit_only=False, nthreads=1):
        """cxx_general_align (src/general_align.cpp:10-62) -> score, edit, ref strings, query strings."""
        sp, so = self._csr(seqs)
        qp, qo = self._csr(quals)
        n = len(so) - 1
        if len(qo) - 1 != n:
            raise OracleError("sequence and quality vectors should have the same length")
        names, err, en = self._enc(encoding)
        score = np.zeros(n, np.float64)
        edit = np.zeros(n, np.int32)
        maxlen = int(np.max(np.diff(so))) if n else 0
        stride = maxlen + len(reference) + 2
        ra = np.zeros((max(n, 1), stride), np.uint8)
        qa = np.zeros((max(n, 1), stride), np.uint8)
        fn = getattr(self.lib, self.p + "_general_align")
        fn.restype = C.c_int
        self._call(fn, C.c_int64(n), sp.ctypes.data_as(C.c_char_p), so.ctypes.data_as(C.c_void_p),
                   qp.ctypes.data_as(C.c_char_p), qo.ctypes.data_as(C.c_void_p),
                   C.c_int(en), names, err.ctypes.data_as(C.c_void_p), C.c_double(go), C.c_double(ge),
                   reference.encode("latin-1"), C.c_int(1 if edit_only else 0),
                   score.ctypes.data_as(C.c_void_p), edit.ctypes.data_as(C.c_void_p),
                   ra.ctypes.data_as(C.c_char_p), qa.ctypes.data_as(C.c_char_p), C.c_int64(stride), C.c_int(nthreads))
        if edit_only:
            return score, edit, [], []
        rs = [bytes(ra[i]).split(b"\0", 1)[0].decode("latin-1") for i in range(n)]
        qs = [bytes(qa[i]).split(b"\0", 1)[0].decode("latin-1") for i in range(n)]
        return score, edit, rs, qs

    def cost_tables(self, encoding):
        names, err, en = self._enc(encoding)
        match = np.zeros((4, en), np.float64)
        mismatch = np.zeros((4, en), np.float64)
        off = C.c_char()
        fn = getattr(self.lib, self.p + "_cost_tables")
        fn.restype = C.c_int
        self._call(fn, C.c_int(en), names, err.ctypes.data_as(C.c_void_p),
                   match.ctypes.data_as(C.c_void_p), mismatch.ctypes.data_as(C.c_void_p), C.byref(off))
        return match, mismatch, off.value
