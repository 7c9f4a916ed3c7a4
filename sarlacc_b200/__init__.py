"""sarlacc_b200: B200-native (sm_100a CUDA) implementation of sarlacc's adaptor-alignment hot path.

Importing this package loads sarlacc_b200/libsarlacc_b200.so; it raises if the library is missing
(there is no CPU fallback).  `native` mirrors the reference's four `.Call` entry points, `api` mirrors
the R drivers around them (adaptorAlign, getAdaptorThresholds, barcodeAlign, ...).
"""
from . import _lib  # noqa: F401  (fails loudly when the native library is absent)
from .reads import ReadSet, read_fastq, read_fastq_condensed, write_fastq  # noqa: F401
from . import native  # noqa: F401
from .native import SarlaccError, phred_encoding, SEQ_ASCII, SEQ_BIOSTRINGS  # noqa: F401
