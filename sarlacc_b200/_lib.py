"""ctypes binding of include/sarlacc_b200.h (the C ABI in sarlacc_b200/libsarlacc_b200.so).

The library is the product: there is no Python or CPU implementation of the alignment behind it.
If the shared object is missing or cannot be loaded this module raises, loudly, at import time.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SARLACC_LIB") or os.path.join(HERE, "libsarlacc_b200.so")   # override: A/B builds of the same library

SEQ_ASCII = 0
SEQ_BIOSTRINGS = 1


class SarlaccError(RuntimeError):
    """An error raised by the native library; the text is the reference's own message where one exists
    (what BEGIN_RCPP/END_RCPP would have turned into an R error)."""


class _Reads(C.Structure):
    _fields_ = [
        ("n", C.c_int64),
        ("seq_encoding", C.c_int),
        ("seq", C.POINTER(C.c_void_p)),
        ("seq_len", C.POINTER(C.c_int32)),
        ("qual", C.POINTER(C.c_void_p)),
        ("qual_len", C.POINTER(C.c_int32)),
        ("seq_pool", C.c_void_p),
        ("seq_off", C.c_void_p),
        ("qual_pool", C.c_void_p),
        ("qual_off", C.c_void_p),
    ]


class _Encoding(C.Structure):
    _fields_ = [("n", C.c_int), ("names", C.POINTER(C.c_char_p)), ("err", C.c_void_p)]


EXPORTS = [
    "sarlacc_last_error", "sarlacc_device_count", "sarlacc_set_devices", "sarlacc_set_host_threads",
    "sarlacc_version", "sarlacc_kernel_launches",
    "sarlacc_adaptor_align", "sarlacc_adaptor_align_score_only", "sarlacc_barcode_align", "sarlacc_general_align",
    "sarlacc_barcode_align_multi", "sarlacc_adaptor_align_windows", "sarlacc_adaptor_align_reads",
    "sarlacc_resident_create", "sarlacc_resident_free", "sarlacc_resident_n", "sarlacc_resident_cells",
    "sarlacc_resident_bytes", "sarlacc_resident_align", "sarlacc_resident_fetch",
    "sarlacc_resident_scores_device", "sarlacc_resident_last_kernel",
    "sarlacc_resident_set_timing", "sarlacc_resident_forward_ms",
    "sarlacc_resident_scrambled", "sarlacc_resident_rows",
    "sarlacc_fastq_open", "sarlacc_fastq_next", "sarlacc_fastq_close",
]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "sarlacc_b200: native library %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C sarlacc_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.sarlacc_last_error.restype = C.c_char_p
    lib.sarlacc_version.restype = C.c_char_p
    lib.sarlacc_kernel_launches.restype = C.c_int64
    lib.sarlacc_kernel_launches.argtypes = [C.c_int]
    lib.sarlacc_resident_create.restype = C.c_void_p
    lib.sarlacc_resident_create.argtypes = [C.POINTER(_Reads), C.POINTER(_Encoding), C.c_int]
    lib.sarlacc_resident_scrambled.restype = C.c_void_p
    lib.sarlacc_resident_scrambled.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int]
    lib.sarlacc_resident_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sarlacc_fastq_open.restype = C.c_void_p
    lib.sarlacc_fastq_open.argtypes = [C.c_char_p]
    lib.sarlacc_fastq_close.restype = None
    lib.sarlacc_fastq_close.argtypes = [C.c_void_p]
    lib.sarlacc_fastq_next.restype = C.c_int64
    lib.sarlacc_fastq_next.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 6
    lib.sarlacc_fastq_next_condensed.restype = C.c_int64
    lib.sarlacc_fastq_next_condensed.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int] + [C.c_void_p] * 7
    lib.sarlacc_resident_free.argtypes = [C.c_void_p]
    lib.sarlacc_resident_free.restype = None
    for name in ("sarlacc_resident_n", "sarlacc_resident_bytes"):
        getattr(lib, name).restype = C.c_int64
        getattr(lib, name).argtypes = [C.c_void_p]
    lib.sarlacc_resident_cells.restype = C.c_int64
    lib.sarlacc_resident_cells.argtypes = [C.c_void_p, C.c_int]
    lib.sarlacc_resident_scores_device.restype = C.c_void_p
    lib.sarlacc_resident_scores_device.argtypes = [C.c_void_p]
    lib.sarlacc_resident_last_kernel.restype = C.c_char_p
    lib.sarlacc_resident_last_kernel.argtypes = [C.c_void_p]
    lib.sarlacc_resident_set_timing.argtypes = [C.c_void_p, C.c_int]
    lib.sarlacc_resident_set_timing.restype = None
    lib.sarlacc_resident_forward_ms.argtypes = [C.c_void_p]
    lib.sarlacc_resident_forward_ms.restype = C.c_double
    lib.sarlacc_resident_align.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_char_p,
                                           C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sarlacc_resident_fetch.argtypes = [C.c_void_p] + [C.c_void_p] * 6
    for name in ("sarlacc_umi_group", "sarlacc_umi_neighbors"):
        getattr(lib, name).restype = C.c_void_p
        getattr(lib, name).argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
    lib.sarlacc_cluster_umis.restype = C.c_void_p
    lib.sarlacc_cluster_umis.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    lib.sarlacc_lists_count.restype = C.c_int64
    lib.sarlacc_lists_count.argtypes = [C.c_void_p]
    lib.sarlacc_lists_values.restype = C.c_int64
    lib.sarlacc_lists_values.argtypes = [C.c_void_p]
    lib.sarlacc_lists_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sarlacc_lists_free.restype = None
    lib.sarlacc_lists_free.argtypes = [C.c_void_p]
    lib.sarlacc_pack_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    lib.sarlacc_pack_bases.restype = C.c_int
    lib.sarlacc_pack_bases.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int]
    lib.sarlacc_chunk_create.restype = C.c_void_p
    lib.sarlacc_chunk_create.argtypes = [C.c_int, C.c_int64, C.c_int, C.POINTER(_Encoding)]
    lib.sarlacc_chunk_free.restype = None
    lib.sarlacc_chunk_free.argtypes = [C.c_void_p]
    lib.sarlacc_chunk_n.restype = C.c_int64
    lib.sarlacc_chunk_n.argtypes = [C.c_void_p]
    lib.sarlacc_chunk_load_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.sarlacc_chunk_load_mock.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_char_p, C.c_char_p, C.c_int,
                                            C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int]
    lib.sarlacc_chunk_adaptor_align.argtypes = ([C.c_void_p, C.c_double, C.c_double, C.c_char_p, C.c_char_p,
                                                 C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64] +
                                                [C.c_void_p] * 12)
    lib.sarlacc_chunk_scrambled_scores.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_char_p, C.c_char_p,
                                                   C.c_uint64, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sarlacc_tied_overlap.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    lib.sarlacc_chunk_sync.argtypes = [C.c_void_p]
    lib.sarlacc_chunk_join.argtypes = [C.c_void_p]
    lib.sarlacc_chunk_stream.restype = C.c_void_p
    lib.sarlacc_chunk_stream.argtypes = [C.c_void_p]
    lib.sarlacc_last_pair_timing.restype = None
    lib.sarlacc_last_pair_timing.argtypes = [C.c_void_p]
    lib.sarlacc_last_pair_upload_bytes.restype = C.c_int64
    lib.sarlacc_last_pair_upload_bytes.argtypes = []
    lib.sarlacc_chunk_rows.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
    lib.sarlacc_chunk_set_timing.restype = None
    lib.sarlacc_chunk_set_timing.argtypes = [C.c_void_p, C.c_int]
    lib.sarlacc_chunk_phase_ms.argtypes = [C.c_void_p, C.c_void_p]
    lib.sarlacc_chunk_last_kernel.restype = C.c_char_p
    lib.sarlacc_chunk_last_kernel.argtypes = [C.c_void_p, C.c_int]
    lib.sarlacc_compute_threshold.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_void_p]
    return lib


lib = _load()


def last_error():
    return lib.sarlacc_last_error().decode("latin-1")


def check(rc):
    if rc != 0:
        raise SarlaccError(last_error())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class ReadsArg:
    """Keeps the numpy buffers behind a sarlacc_reads struct alive."""

    def __init__(self, seq_pool, seq_off, qual_pool, qual_off, seq_encoding=SEQ_ASCII, views=False):
        self.seq_pool = np.ascontiguousarray(seq_pool, dtype=np.uint8)
        self.qual_pool = np.ascontiguousarray(qual_pool, dtype=np.uint8)
        if self.seq_pool.size == 0:
            self.seq_pool = np.zeros(1, np.uint8)
        if self.qual_pool.size == 0:
            self.qual_pool = np.zeros(1, np.uint8)
        self.seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
        self.qual_off = np.ascontiguousarray(qual_off, dtype=np.int64)
        self.n = len(self.seq_off) - 1
        self.struct = _Reads()
        self.struct.n = self.n
        self.struct.seq_encoding = seq_encoding
        if views:
            # the XStringSet-holder layout: one (pointer, length) pair per element
            sbase = self.seq_pool.ctypes.data
            qbase = self.qual_pool.ctypes.data
            self._sp = (sbase + self.seq_off[:-1]).astype(np.uint64)
            self._qp = (qbase + self.qual_off[:-1]).astype(np.uint64)
            self._sl = np.diff(self.seq_off).astype(np.int32)
            self._ql = np.diff(self.qual_off).astype(np.int32)
            if self.n == 0:
                self._sp = np.zeros(1, np.uint64); self._qp = np.zeros(1, np.uint64)
                self._sl = np.zeros(1, np.int32); self._ql = np.zeros(1, np.int32)
            self.struct.seq = self._sp.ctypes.data_as(C.POINTER(C.c_void_p))
            self.struct.qual = self._qp.ctypes.data_as(C.POINTER(C.c_void_p))
            self.struct.seq_len = self._sl.ctypes.data_as(C.POINTER(C.c_int32))
            self.struct.qual_len = self._ql.ctypes.data_as(C.POINTER(C.c_int32))
        else:
            self.struct.seq_pool = self.seq_pool.ctypes.data
            self.struct.seq_off = self.seq_off.ctypes.data
            self.struct.qual_pool = self.qual_pool.ctypes.data
            self.struct.qual_off = self.qual_off.ctypes.data

    def ref(self):
        return C.byref(self.struct)


class EncodingArg:
    def __init__(self, names, err):
        self.err = np.ascontiguousarray(err, dtype=np.float64)
        self.struct = _Encoding()
        self.struct.n = len(self.err)
        if names is None:
            self._names = None
            self.struct.names = None
        else:
            enc = [x.encode("latin-1") if isinstance(x, str) else bytes(x) for x in names]
            self._names = (C.c_char_p * max(len(enc), 1))(*enc)
            self.struct.names = C.cast(self._names, C.POINTER(C.c_char_p))
        self.struct.err = self.err.ctypes.data

    def ref(self):
        return C.byref(self.struct)
