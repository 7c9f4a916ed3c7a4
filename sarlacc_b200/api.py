"""Host-side mirror of the reference's R drivers around the adaptor-alignment hot path.

Same names, arguments, defaults, return columns and quirks as the R functions they mirror, so that the
parity tests read like the reference's own tests:

    adaptorAlign           R/adaptorAlign.R:7-78
    getAdaptorThresholds   R/getAdaptorThresholds.R:6-66
    barcodeAlign           R/barcodeAlign.R:4-40
    tuneAlignment          R/tuneAlignment.R:6-76          (caller of the same entry points)
    extractSubseq          R/extractSubseq.R:5-117         (caller; re-aligns and re-checks the stored scores)
    qualityAlign           R/qualityAlign.R                (general_align wrapper)
    umiGroup, qualityMask  R/umiGroup.R:2-23, R/qualityMask.R:5-15   (neighbour search on the device)
    helpers: _setup_subseqs (:136-143), _get_front_and_back (:86-95), _resolve_strand (:112-122),
             _parallelize (:126-134), _align_and_extract (:150-176), _align_AA_internal (:178-199),
             _scramble_input (getAdaptorThresholds.R:68-92), _compute_threshold (:94-103),
             _tied_overlap (tuneAlignment.R:78-86), _create_encoding_vector (qualityMask.R:19-27)

All alignment arithmetic happens in the CUDA library (sarlacc_b200.native); this module only does what
the R code does around the `.Call`s.  R data structures map to: QualityScaledDNAStringSet -> ReadSet,
DataFrame -> Frame (ordered columns + metadata + rownames), NA_integer_ -> 0 for barcode ids.
"""
import re

import numpy as np

from . import native
from .reads import ReadSet, read_fastq, read_fastq_condensed


# --------------------------------------------------------------------------------------------------
# A minimal DataFrame
# --------------------------------------------------------------------------------------------------
class Frame:
    """Ordered named columns of equal length (numpy arrays, ReadSets or nested Frames) + metadata + rownames."""

    def __init__(self, columns=None, nrows=None, rownames=None, metadata=None):
        self.columns = dict(columns or {})
        self._nrows = nrows
        self.rownames = rownames
        self.metadata = dict(metadata or {})

    def __len__(self):
        if self._nrows is not None:
            return self._nrows
        for v in self.columns.values():
            return len(v)
        return 0

    def __getitem__(self, key):
        return self.columns[key]

    def __setitem__(self, key, value):
        self.columns[key] = value

    def __contains__(self, key):
        return key in self.columns

    def __getattr__(self, key):
        cols = self.__dict__.get("columns", {})
        if key in cols:
            return cols[key]
        raise AttributeError(key)

    def names(self):
        return list(self.columns)

    def assign_rows(self, mask, other):
        """self[mask,] <- other[mask,] for every column (R/adaptorAlign.R:195-196), nested frames included."""
        mask = np.asarray(mask, dtype=bool)
        for k, v in self.columns.items():
            o = other.columns[k]
            if isinstance(v, Frame):
                v.assign_rows(mask, o)
            elif isinstance(v, ReadSet):
                self.columns[k] = _readset_where(mask, o, v)
            else:
                v = v.copy()
                v[mask] = o[mask]
                self.columns[k] = v

    @staticmethod
    def rbind(frames):
        frames = list(frames)
        first = frames[0]
        out = Frame(nrows=sum(len(f) for f in frames), metadata=first.metadata)
        for k, v in first.columns.items():
            parts = [f.columns[k] for f in frames]
            if isinstance(v, Frame):
                out.columns[k] = Frame.rbind(parts)
            elif isinstance(v, ReadSet):
                out.columns[k] = ReadSet.concat(parts)
            else:
                out.columns[k] = np.concatenate(parts)
        if all(f.rownames is not None for f in frames):
            out.rownames = [x for f in frames for x in f.rownames]
        return out


def _readset_where(mask, a, b):
    """Element-wise ifelse(mask, a, b) for two ReadSets of equal length."""
    n = len(a)
    both = ReadSet.concat([a, b])
    idx = np.where(mask, np.arange(n), n + np.arange(n))
    return both[idx]


# --------------------------------------------------------------------------------------------------
# helpers (dot-prefixed internals of the R package)
# --------------------------------------------------------------------------------------------------
def _qual2class(qual_type):
    # R/adaptorAlign.R:97-99
    return qual_type[:1].upper() + qual_type[1:] + "Quality"


_QUAL_ENCODINGS = {
    # Biostrings::encoding(): offset character and score range of each quality class
    "PhredQuality": (33, 0, 93, "phred"),
    "IlluminaQuality": (64, 0, 62, "phred"),
    "SolexaQuality": (64, -5, 62, "solexa"),      # ';' (59) .. '~' (126): offset 64, scores -5..62
}


def _create_encoding_vector(qual_class="PhredQuality"):
    """R/qualityMask.R:19-27: names = the encoding characters, values = their error probabilities
    (Phred: 10^(-q/10); Solexa: 1/(1+10^(q/10)) ... expressed as an error probability)."""
    offset, lo, hi, kind = _QUAL_ENCODINGS[qual_class]
    qs = np.arange(lo, hi + 1)
    names = [chr(offset + int(q)) for q in qs]
    # scalar libm pow per entry (not numpy's vectorised pow, which may differ in the last bit): the table must be
    # the same numbers whoever builds it, since scores are compared bit for bit
    if kind == "phred":
        err = np.array([10.0 ** (-int(q) / 10.0) for q in qs], dtype=np.float64)
    else:
        err = np.array([1.0 - 1.0 / (1.0 + 10.0 ** (-int(q) / 10.0)) for q in qs], dtype=np.float64)
    return names, err


def _setup_subseqs(adaptor):
    """R/adaptorAlign.R:136-143: 1-based starts/ends of the runs of non-ACGT characters."""
    starts, ends = [], []
    for m in re.finditer("[^ACTG]+", adaptor):
        starts.append(m.start() + 1)
        ends.append(m.end())
    return {"starts": np.array(starts, dtype=np.int32), "ends": np.array(ends, dtype=np.int32)}


def _get_front_and_back(reads, tolerance):
    """R/adaptorAlign.R:86-95: first `tolerance` bases, and the reverse complement of the last `tolerance` bases."""
    w = reads.width()
    tol = np.minimum(np.int64(tolerance), w)
    front = reads.subseq(start=np.ones(len(reads), np.int64), width=tol)
    back = reads.subseq(end=w, width=tol).reverse_complement()
    back.names = reads.names
    return {"front": front, "back": back}


def _resolve_strand(start_score, end_score, rc_start_score, rc_end_score):
    """R/adaptorAlign.R:112-122."""
    fscore = np.maximum(start_score, 0) + np.maximum(end_score, 0)
    rscore = np.maximum(rc_start_score, 0) + np.maximum(rc_end_score, 0)
    is_reverse = fscore < rscore
    return {"reversed": is_reverse, "scores": np.where(is_reverse, rscore, fscore)}


def _parallelize(n, n_workers):
    """R/adaptorAlign.R:126-134: contiguous split of seq_len(n) into n_workers chunks; returns index arrays.
    (The GPU library shards by the same rule internally; this is kept for API parity and the gloo tests.)"""
    if n == 0:
        return []
    bounds = np.linspace(1, n, n_workers + 1)[:-1]
    ids = np.searchsorted(bounds, np.arange(1, n + 1), side="right")
    return [np.nonzero(ids == k)[0] for k in np.unique(ids)]


def _align_and_extract(adaptor, reads, gap_opening, gap_extension, subseq_starts, subseq_ends, encoding=None):
    """R/adaptorAlign.R:150-176."""
    enc = encoding or native.phred_encoding()
    out = native.adaptor_align(reads, enc, gap_opening, gap_extension, adaptor,
                               np.asarray(subseq_starts, dtype=np.int32) - 1, subseq_ends)
    output = Frame({"score": out[0], "start": out[1], "end": out[2]})
    segments = Frame(nrows=len(reads))
    for i in range(len(out[3])):
        segments["Sub%d" % (i + 1)] = reads.subseq(start=out[3][i], width=out[4][i])
    output["subseq"] = segments
    return output


def _align_AA_internal(reads, adaptor1, adaptor2, tolerance, subseq1, subseq2, gap_opening, gap_extension, encoding=None):
    """R/adaptorAlign.R:178-199."""
    w = _get_front_and_back(reads, tolerance)
    args = dict(gap_opening=gap_opening, gap_extension=gap_extension, encoding=encoding)
    cur_starts = _align_and_extract(adaptor1, w["front"], subseq_starts=subseq1["starts"], subseq_ends=subseq1["ends"], **args)
    cur_ends = _align_and_extract(adaptor2, w["back"], subseq_starts=subseq2["starts"], subseq_ends=subseq2["ends"], **args)
    cur_rc_starts = _align_and_extract(adaptor1, w["back"], subseq_starts=subseq1["starts"], subseq_ends=subseq1["ends"], **args)
    cur_rc_ends = _align_and_extract(adaptor2, w["front"], subseq_starts=subseq2["starts"], subseq_ends=subseq2["ends"], **args)
    strand = _resolve_strand(cur_starts["score"], cur_ends["score"], cur_rc_starts["score"], cur_rc_ends["score"])
    is_reverse = strand["reversed"]
    cur_starts.assign_rows(is_reverse, cur_rc_starts)
    cur_ends.assign_rows(is_reverse, cur_rc_ends)
    return {"names": reads.names, "width": reads.width(), "start": cur_starts, "end": cur_ends, "reversed": is_reverse}


def _window_subseq(reads, use_back, start, width):
    """subseq(window, start, width) without materialising the windows of .get_front_and_back: rows with use_back
    take it from the reverse-complemented tail of the read (window position p <-> read position W - p + 1)."""
    W = reads.width()
    start = np.asarray(start, dtype=np.int64)
    width = np.asarray(width, dtype=np.int64)
    fwd = reads.subseq(start=np.where(use_back, 1, start), width=np.where(use_back, 0, width))
    bstart = np.where(use_back, W - start - width + 2, 1)
    bwd = reads.subseq(start=bstart, width=np.where(use_back, width, 0)).reverse_complement()
    return _readset_where(use_back, bwd, fwd)


def _align_AA_internal_fused(reads, adaptor1, adaptor2, tolerance, subseq1, subseq2, gap_opening, gap_extension, encoding=None):
    """_align_AA_internal with window cutting, the four .Calls, .resolve_strand and the row selection done in one
    library call (sarlacc_adaptor_align_reads).  Same return value, except that adaptor2's start/end are already in
    read coordinates (flagged by "flipped")."""
    enc = encoding or native.phred_encoding()
    if len(adaptor1) == 0 or len(adaptor2) == 0 or len(reads) == 0 or int(tolerance) < 1:
        # degenerate inputs the reference accepts (tolerance 0: empty windows, the score is the row-0 chain) take its own route
        return _align_AA_internal(reads, adaptor1, adaptor2, tolerance, subseq1, subseq2, gap_opening, gap_extension, encoding)
    width, rev, r1, r2 = native.adaptor_align_reads(
        reads, tolerance, enc, gap_opening, gap_extension, adaptor1, adaptor2,
        (np.asarray(subseq1["starts"], dtype=np.int32) - 1, subseq1["ends"]),
        (np.asarray(subseq2["starts"], dtype=np.int32) - 1, subseq2["ends"]))
    frames = []
    # the sections were located on the window that was kept: front for adaptor1 / back for adaptor2, swapped if reversed
    for out, use_back in ((r1, rev), (r2, ~rev)):
        f = Frame({"score": out[0], "start": out[1], "end": out[2]})
        seg = Frame(nrows=len(reads))
        for i in range(len(out[3])):
            seg["Sub%d" % (i + 1)] = _window_subseq(reads, use_back, out[3][i], out[4][i])
        f["subseq"] = seg
        frames.append(f)
    return {"names": reads.names, "width": width.astype(np.int64), "start": frames[0], "end": frames[1], "reversed": rev,
            "flipped": True}


def _prefetch(gen, depth=2):
    """Runs a generator in a background thread, `depth` items ahead: the next chunk is parsed (in C, GIL released)
    while the device works on the current one.  Closing the consumer early (tuneAlignment takes one chunk) stops
    the producer and closes the underlying generator."""
    import queue
    import threading
    q = queue.Queue(maxsize=depth)
    done = object()
    stop = threading.Event()

    def put(item):
        while not stop.is_set():
            try:
                q.put(item, timeout=0.05)
                return True
            except queue.Full:
                continue
        return False

    def work():
        try:
            for item in gen:
                if not put(item):
                    break
            else:
                put(done)
        except BaseException as e:      # surfaced in the consumer
            put(e)
        finally:
            if hasattr(gen, "close"):
                gen.close()

    t = threading.Thread(target=work, daemon=True)
    t.start()
    try:
        while True:
            item = q.get()
            if item is done:
                return
            if isinstance(item, BaseException):
                raise item
            yield item
    finally:
        stop.set()
        t.join(timeout=5)


def _stream(source, number, keep=None):
    """FastqStreamer(filepath, n=number) + yield (R/adaptorAlign.R:26,36), or chunks of an in-memory ReadSet.
    Yields (reads, true widths or None).  With `keep` set, plain-text FASTQ goes through the parallel condensed ingest:
    only the first and last `keep` bases of every read are materialised (all the alignment looks at), widths carry
    the real read lengths."""
    number = int(number)
    if isinstance(source, ReadSet):
        n = len(source)
        for lo in range(0, n, number):
            idx = np.arange(lo, min(n, lo + number))
            yield source[idx], None
    elif keep is not None and int(keep) >= 1:
        yield from _prefetch(read_fastq_condensed(source, keep, number))
    else:
        for reads in read_fastq(source, number):
            yield reads, None


# --------------------------------------------------------------------------------------------------
# exported functions
# --------------------------------------------------------------------------------------------------
def adaptorAlign(adaptor1, adaptor2, filepath, tolerance=250, gapOpening=5, gapExtension=1,
                 qual_type="phred", number=1e5, fused=True):
    """R/adaptorAlign.R:7-78.  `filepath` is a FASTQ path or an in-memory ReadSet (names required for
    getAdaptorThresholds).  Returns a Frame with read.width, adaptor1, adaptor2, reversed and the same metadata.
    fused=True runs .align_AA_internal's four alignments + strand resolution in one device pass
    (sarlacc_adaptor_align_windows); fused=False issues the reference's four .Calls.  Results are identical."""
    adaptor1 = str(adaptor1).upper()
    adaptor2 = str(adaptor2).upper()
    if qual_type not in ("phred", "solexa", "illumina"):
        raise ValueError("'arg' should be one of 'phred', 'solexa', 'illumina'")
    enc = _create_encoding_vector(_qual2class(qual_type))
    all_args = dict(adaptor1=adaptor1, adaptor2=adaptor2, tolerance=tolerance,
                    subseq1=_setup_subseqs(adaptor1), subseq2=_setup_subseqs(adaptor2),
                    gap_opening=gapOpening, gap_extension=gapExtension, encoding=enc)
    names, widths, starts, ends, revs = [], [], [], [], []
    internal = _align_AA_internal_fused if fused else _align_AA_internal
    for reads, true_width in _stream(filepath, number, keep=int(tolerance)):
        out = internal(reads, **all_args)
        if out.get("flipped"):
            # undo the library's adaptor2 flip so that the single flip below (R/adaptorAlign.R:66-71) applies to every chunk
            w = out["width"]
            out["end"]["start"] = (w - out["end"]["start"] + 1).astype(np.int32)
            out["end"]["end"] = (w - out["end"]["end"] + 1).astype(np.int32)
        names.append(out["names"] if out["names"] is not None else [None] * len(reads))
        # condensed ingest: the windows are the real read's windows, the width is not
        widths.append(out["width"] if true_width is None else true_width.astype(np.int64))
        starts.append(out["start"])
        ends.append(out["end"])
        revs.append(out["reversed"])
    if not starts:
        out = _align_AA_internal(ReadSet.empty(), **all_args)   # guarantee some value is returned (:47-55)
        names, widths, starts, ends, revs = [[]], [out["width"]], [out["start"]], [out["end"]], [out["reversed"]]
    align_start = Frame.rbind(starts)
    align_end = Frame.rbind(ends)
    details = {"gapOpening": gapOpening, "gapExtension": gapExtension}
    align_start.metadata = dict(sequence=adaptor1, **details)
    align_end.metadata = dict(sequence=adaptor2, **details)
    all_widths = np.concatenate(widths).astype(np.int64)
    # Adjusting the reverse coordinates for the read length (:66-71)
    align_end["start"] = (all_widths - align_end["start"] + 1).astype(np.int64)
    align_end["end"] = (all_widths - align_end["end"] + 1).astype(np.int64)
    all_names = [x for chunk in names for x in chunk]
    align_start.rownames = align_end.rownames = all_names
    output = Frame({"read.width": all_widths, "adaptor1": align_start, "adaptor2": align_end,
                    "reversed": np.concatenate(revs)}, rownames=all_names,
                   metadata={"filepath": filepath, "qual.type": qual_type, "tolerance": tolerance})
    return output


_U32 = np.uint32


def _hash32(x):
    """kernels.cu: hash32 (two 32-bit multiplies), on uint32 arrays."""
    x = np.asarray(x, dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        x ^= x >> _U32(16)
        x *= _U32(0x21F0AAAD)
        x ^= x >> _U32(15)
        x *= _U32(0x735A2D97)
        x ^= x >> _U32(15)
    return x


def _stream_key(seed, index, field):
    """kernels.cu: stream_key -- the 32-bit key of the stream (seed, global read index, field); index may be an array."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    index = np.asarray(index, dtype=np.uint64)
    k = _hash32(_U32((seed & 0xFFFFFFFF) ^ 0x243F6A88))
    k = _hash32(k ^ _U32(seed >> 32))
    k = _hash32(k ^ (index & np.uint64(0xFFFFFFFF)).astype(np.uint32))
    k = _hash32(k ^ (index >> np.uint64(32)).astype(np.uint32))
    with np.errstate(over="ignore"):
        f = _U32((int(field) * 0x85EBCA6B + 0xC2B2AE35) & 0xFFFFFFFF)
    return _hash32(k ^ f)


def _stream_word(key, p):
    """kernels.cu: stream_word -- word p of the stream with that key (broadcasts)."""
    with np.errstate(over="ignore"):
        return _hash32(np.asarray(key, np.uint32) ^ (np.asarray(p, np.uint32) * _U32(0x9E3779B1) + _U32(0x7F4A7C15)))


def _scramble_by_index(seqs, seed, read_index, stream):
    """R/getAdaptorThresholds.R:68-92: one uniform random permutation per sequence, applied to bases and qualities
    alike.  R's sample() stream cannot be reproduced outside R; here it is a Fisher-Yates shuffle -- for i = len-1 .. 1:
    j = floor(word_i * (i + 1) / 2^32), swap(i, j) -- driven by the counter-based stream of (seed, global read index,
    field 16 + stream), so it does not depend on chunking, sharding or device count.  kernels.cu: scramble_rows_fy is the
    same function on the device."""
    n = len(seqs)
    w = seqs.width().astype(np.int64)
    maxw = int(w.max()) if n else 0
    key = _stream_key(seed, np.asarray(read_index, dtype=np.uint64), 16 + int(stream))
    perm = np.tile(np.arange(max(maxw, 1), dtype=np.int64), (n, 1))        # perm[r, i] = source position of output i
    rows = np.arange(n)
    for i in range(maxw - 1, 0, -1):
        live = w > i
        if not live.any():
            continue
        j = ((_stream_word(key, i).astype(np.uint64) * np.uint64(i + 1)) >> np.uint64(32)).astype(np.int64)
        r = rows[live]
        jj = j[live]
        a, b = perm[r, i].copy(), perm[r, jj].copy()
        perm[r, i] = b
        perm[r, jj] = a
    off = np.zeros(n, dtype=np.int64)
    if n:
        np.cumsum(w[:-1], out=off[1:])
    total = int(w.sum())
    pos = np.arange(total, dtype=np.int64) - np.repeat(off, w)
    src = np.repeat(seqs.seq_off[:-1] if n else np.zeros(0, np.int64), w) + perm[np.repeat(rows, w), pos]
    sp = seqs.seq_pool[src]
    qp = None
    if seqs.has_quality:
        qsrc = np.repeat(seqs.qual_off[:-1] if n else np.zeros(0, np.int64), w) + perm[np.repeat(rows, w), pos]
        qp = seqs.qual_pool[qsrc]
    o = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(w, out=o[1:])
    return ReadSet(sp, o, qp, o if qp is not None else None, seqs.names)


def _scramble_input(seqs, has_qual=True, seed=0, first_index=0, stream=0):
    """_scramble_by_index for reads first_index, first_index + 1, ... (has_qual=False drops the qualities)."""
    out = _scramble_by_index(seqs, seed, np.arange(first_index, first_index + len(seqs), dtype=np.uint64), stream)
    if not has_qual and out.has_quality:
        out = ReadSet(out.seq_pool, out.seq_off, None, None, out.names)
    return out


def _compute_threshold(real, scrambled, error):
    """R/getAdaptorThresholds.R:94-103, IEEE semantics included (x/0 -> Inf/NaN, dropped by which())."""
    real = np.sort(np.asarray(real, dtype=np.float64))
    scrambled = np.sort(np.asarray(scrambled, dtype=np.float64))
    found = np.searchsorted(scrambled, real, side="right")          # findInterval(real, scrambled)
    with np.errstate(divide="ignore", invalid="ignore"):
        fdr = (len(scrambled) - found) / (len(real) - np.arange(1, len(real) + 1, dtype=np.float64))
    ok = np.nonzero(fdr <= error)[0]
    if len(ok) == 0:
        return float("nan")                                         # real[min(integer(0))] -> NA
    return float(real[ok[0]])


def _get_alignment_scores(reads_start, reads_end, adaptor1, adaptor2, gap_opening, gap_extension, encoding=None):
    """R/tuneAlignment.R:99-112."""
    enc = encoding or native.phred_encoding()

    def fun(r, a):
        return native.adaptor_align_score_only(r, enc, gap_opening, gap_extension, a)

    return {"START": fun(reads_start, adaptor1), "END": fun(reads_end, adaptor2),
            "RSTART": fun(reads_end, adaptor1), "REND": fun(reads_start, adaptor2)}


def _align_AT_internal(reads, adaptor1, adaptor2, tolerance, gap_opening, gap_extension, encoding=None, seed=0, first_index=0):
    """R/getAdaptorThresholds.R:105-128."""
    w = _get_front_and_back(reads, tolerance)
    scr_start = _scramble_input(w["front"], True, seed, first_index, 0)
    scr_end = _scramble_input(w["back"], True, seed, first_index, 1)
    sc = _get_alignment_scores(scr_start, scr_end, adaptor1, adaptor2, gap_opening, gap_extension, encoding)
    is_reverse = _resolve_strand(sc["START"], sc["END"], sc["RSTART"], sc["REND"])["reversed"]
    return {"adaptor1": np.where(is_reverse, sc["RSTART"], sc["START"]),
            "adaptor2": np.where(is_reverse, sc["REND"], sc["END"])}


def getAdaptorThresholds(aligned, error=0.01, number=1e5, seed=0, device_scramble=True):
    """R/getAdaptorThresholds.R:6-66.  device_scramble=True permutes the windows on the GPU; False builds the same
    keyed permutation with numpy and goes through the reference's four score-only .Calls.  Identical results."""
    go = aligned["adaptor1"].metadata["gapOpening"]
    ge = aligned["adaptor1"].metadata["gapExtension"]
    adaptor1 = aligned["adaptor1"].metadata["sequence"]
    adaptor2 = aligned["adaptor2"].metadata["sequence"]
    tolerance = aligned.metadata["tolerance"]
    filepath = aligned.metadata["filepath"]
    enc = _create_encoding_vector(_qual2class(aligned.metadata["qual.type"]))
    wanted = {nm: k for k, nm in enumerate(aligned.rownames)}
    scr1, scr2, used = [], [], []
    seen = 0
    chunk = None      # device path: one chunk object re-loaded per FASTQ chunk (.align_AT_internal = sarlacc_chunk_scrambled_scores)
    try:
        for reads, _ in _stream(filepath, number, keep=int(tolerance)):     # thresholds only look at the windows
            keep = np.array([nm in wanted for nm in reads.names], dtype=bool)
            first_index = seen
            seen += len(reads)
            idx = np.nonzero(keep)[0]
            sub = reads[idx]
            used.extend(sub.names)
            if len(sub) == 0:
                continue
            # the scramble is keyed by the read's position in the file, so dropping reads does not shift others
            if device_scramble and int(tolerance) >= 1:
                if chunk is None or chunk.capacity < len(sub):
                    if chunk is not None:
                        chunk.close()
                    chunk = native.Chunk(max(len(sub), int(number)), int(tolerance), enc)
                chunk.load_reads(sub, int(tolerance))
                a1s, a2s = chunk.scrambled_scores(go, ge, adaptor1, adaptor2, seed=seed, read_index=first_index + idx)
                scr1.append(a1s)
                scr2.append(a2s)
                continue
            w = _get_front_and_back(sub, tolerance)
            scr_start = _scramble_by_index(w["front"], seed, first_index + idx, 0)
            scr_end = _scramble_by_index(w["back"], seed, first_index + idx, 1)
            sc = _get_alignment_scores(scr_start, scr_end, adaptor1, adaptor2, go, ge, enc)
            is_reverse = _resolve_strand(sc["START"], sc["END"], sc["RSTART"], sc["REND"])["reversed"]
            scr1.append(np.where(is_reverse, sc["RSTART"], sc["START"]))
            scr2.append(np.where(is_reverse, sc["REND"], sc["END"]))
    finally:
        if chunk is not None:
            chunk.close()
    scram1 = np.concatenate(scr1) if scr1 else np.zeros(0)
    scram2 = np.concatenate(scr2) if scr2 else np.zeros(0)
    m = np.array([wanted[nm] for nm in used], dtype=np.int64)
    real1 = aligned["adaptor1"]["score"][m]
    real2 = aligned["adaptor2"]["score"][m]
    select = native.compute_threshold if (device_scramble and len(real1)) else _compute_threshold     # same numbers (tests)
    return {"threshold1": select(real1, scram1, error),
            "threshold2": select(real2, scram2, error),
            "scores1": {"reads": real1, "scrambled": scram1},
            "scores2": {"reads": real2, "scrambled": scram2}}


def barcodeAlign(sequences, barcodes, gapOpening=5, gapExtension=1, qual_type="phred"):
    """R/barcodeAlign.R:4-40, with the per-barcode loop fused into one device pass.  Quirk kept: the barcodes are
    passed to the aligner as given (the reference computes toupper() at :21 but passes barcodes[b] at :23)."""
    enc = _create_encoding_vector(_qual2class(qual_type))
    barcodes = [str(b) for b in barcodes]
    if len({len(b) for b in barcodes}) <= 1:
        bid, best, nxt = native.barcode_align_multi(sequences, enc, gapOpening, gapExtension, barcodes)
    else:
        # barcodes of different lengths cannot share one pass; fall back to the reference's own loop (:20-35)
        n = len(sequences)
        best = np.full(n, -np.inf)
        nxt = np.full(n, -np.inf)
        bid = np.zeros(n, np.int32)
        for b, bc in enumerate(barcodes):
            scores = native.barcode_align(sequences, enc, gapOpening, gapExtension, bc)
            keep = scores > best
            second = ~keep & (scores > nxt)
            bid[keep] = b + 1
            nxt[keep] = best[keep]
            best[keep] = scores[keep]
            nxt[second] = scores[second]
    with np.errstate(invalid="ignore"):
        gap = best - nxt
    return Frame({"barcode": bid, "score": best, "gap": gap},
                 metadata={"gapOpening": gapOpening, "gapExtension": gapExtension, "barcodes": barcodes})


def _tied_overlap(real, fake):
    """R/tuneAlignment.R:78-86."""
    real = np.asarray(real, dtype=np.float64)
    fake = np.sort(np.asarray(fake, dtype=np.float64))
    upper = np.searchsorted(fake, real, side="right")
    lower = np.searchsorted(fake, real, side="left")
    return float(np.sum((upper + lower) / 2.0) / (len(real) * len(fake)))


def _sample_reads(source, number, tolerance, seed):
    """FastqSampler(filepath, number) + yield (R/tuneAlignment.R:21): a uniform random sample of `number` reads, in file
    order.  R's RNG stream cannot be reproduced outside R; the sample here is reservoir sampling (algorithm R) driven by
    numpy's Philox generator keyed by `seed` -- one pass over the stream, like FastqSampler."""
    number = int(number)
    rng = np.random.Generator(np.random.Philox(key=int(seed)))
    kept = None         # ReadSet of at most `number` reads
    order = None        # their positions in the stream
    seen = 0
    for chunk, _ in _stream(source, max(number, 100000), keep=int(tolerance) if int(tolerance) >= 1 else None):
        m = len(chunk)
        pos = seen + np.arange(m)
        if kept is None:
            take = min(number, m)
            kept, order = chunk[np.arange(take)], pos[:take].copy()
            rest = np.arange(take, m)
        else:
            rest = np.arange(m)
            if len(kept) < number:
                take = min(number - len(kept), m)
                kept = ReadSet.concat([kept, chunk[np.arange(take)]])
                order = np.concatenate([order, pos[:take]])
                rest = np.arange(take, m)
        if len(rest):
            # element at stream position t replaces a uniformly chosen slot with probability number / (t + 1)
            j = (rng.random(len(rest)) * (pos[rest] + 1)).astype(np.int64)
            hit = np.nonzero(j < number)[0]
            if len(hit):
                last = {}
                for h in hit:                       # later replacements of the same slot win
                    last[int(j[h])] = int(rest[h])
                slots = np.array(sorted(last), dtype=np.int64)
                src = np.array([last[int(k)] for k in slots], dtype=np.int64)
                idx = np.arange(len(kept))
                both = ReadSet.concat([kept, chunk[src]])
                idx[slots] = len(kept) + np.arange(len(src))
                kept = both[idx]
                order[slots] = pos[src]
        seen += m
    if kept is None:
        return ReadSet.empty()
    return kept[np.argsort(order, kind="stable")]


def tuneAlignment(adaptor1, adaptor2, filepath, tolerance=200, number=10000, gapOp_range=(4, 10), gapExt_range=(1, 5),
                  qual_type="phred", seed=0, sample="random"):
    """R/tuneAlignment.R:6-76.  sample="random": a uniform sample of `number` reads like FastqSampler's (our own RNG, see
    _sample_reads); sample="first": the first `number` reads of the stream.  The sampled windows are loaded into one
    device-resident chunk and scrambled once; all 35 (gapOpening, gapExtension) points x {reads, scrambled} x 4
    alignments are enqueued back to back with the scores kept on the device, .tied_overlap runs there too, and only the
    winning pair's score vectors come back."""
    adaptor1 = str(adaptor1).upper()
    adaptor2 = str(adaptor2).upper()
    enc = _create_encoding_vector(_qual2class(qual_type))
    if sample == "random":
        reads = _sample_reads(filepath, number, tolerance, seed)
    else:
        reads = None
        for chunk, _ in _stream(filepath, number, keep=int(tolerance) if int(tolerance) >= 1 else None):
            reads = chunk
            break
    if reads is None or len(reads) == 0:
        return {"parameters": {"gapOpening": None, "gapExtension": None},
                "scores": {"reads": np.zeros(0), "scrambled": np.zeros(0)}}
    go_r = np.maximum.accumulate(np.asarray(gapOp_range, dtype=int))
    ge_r = np.maximum.accumulate(np.asarray(gapExt_range, dtype=int))
    grid = [(go, ge) for go in range(int(go_r[0]), int(go_r[1]) + 1) for ge in range(int(ge_r[0]), int(ge_r[1]) + 1)]
    if int(tolerance) < 1 or len(adaptor1) == 0 or len(adaptor2) == 0:
        return _tune_alignment_host(reads, adaptor1, adaptor2, tolerance, grid, enc, seed)
    import torch
    n = len(reads)
    ch = native.Chunk(n, int(tolerance), enc)
    try:
        ch.load_reads(reads, int(tolerance))
        dev = torch.device("cuda", ch.device)
        real = torch.empty((len(grid), n), dtype=torch.float64, device=dev)
        fake = torch.empty((len(grid), n), dtype=torch.float64, device=dev)
        for k, (go, ge) in enumerate(grid):
            ch.scrambled_scores(go, ge, adaptor1, adaptor2, scramble=False, strand_score=real[k].data_ptr())
            ch.scrambled_scores(go, ge, adaptor1, adaptor2, seed=seed, first_index=0, scramble=True if k == 0 else "reuse",
                                strand_score=fake[k].data_ptr())
        ch.sync()
        max_score, final = 0.0, None
        for k, (go, ge) in enumerate(grid):
            cur = native.tied_overlap((real[k].data_ptr(), n), (fake[k].data_ptr(), n), device=ch.device)
            if max_score < cur:
                max_score, final = cur, k
        if final is None:
            return {"parameters": {"gapOpening": None, "gapExtension": None}, "scores": {"reads": None, "scrambled": None}}
        return {"parameters": {"gapOpening": grid[final][0], "gapExtension": grid[final][1]},
                "scores": {"reads": real[final].cpu().numpy(), "scrambled": fake[final].cpu().numpy()}}
    finally:
        ch.close()


def _tune_alignment_host(reads, adaptor1, adaptor2, tolerance, grid, enc, seed):
    """The reference's own loop (R/tuneAlignment.R:54-72) over the four score-only calls: degenerate inputs
    (tolerance 0, an empty adaptor) that the chunk engine does not take."""
    w = _get_front_and_back(reads, tolerance)
    scr_start = _scramble_input(w["front"], True, seed, 0, 0)
    scr_end = _scramble_input(w["back"], True, seed, 0, 1)
    max_score, final = 0.0, None
    for go, ge in grid:
        a = _get_alignment_scores(w["front"], w["back"], adaptor1, adaptor2, go, ge, enc)
        b = _get_alignment_scores(scr_start, scr_end, adaptor1, adaptor2, go, ge, enc)
        read_scores = _resolve_strand(a["START"], a["END"], a["RSTART"], a["REND"])["scores"]
        scr_scores = _resolve_strand(b["START"], b["END"], b["RSTART"], b["REND"])["scores"]
        cur = _tied_overlap(read_scores, scr_scores)
        if max_score < cur:
            max_score, final = cur, (go, ge, read_scores, scr_scores)
    if final is None:
        return {"parameters": {"gapOpening": None, "gapExtension": None}, "scores": {"reads": None, "scrambled": None}}
    return {"parameters": {"gapOpening": final[0], "gapExtension": final[1]},
            "scores": {"reads": final[2], "scrambled": final[3]}}


def _extract_internal(reads, flipped, adaptor1, adaptor2, tolerance, subseq1, subseq2, gap_opening, gap_extension, encoding=None):
    """R/extractSubseq.R:89-117: re-align only the window the stored orientation points at."""
    w = _get_front_and_back(reads, tolerance)
    flipped = np.asarray(flipped, dtype=bool)
    actual_starts = _readset_where(flipped, w["back"], w["front"])
    actual_ends = _readset_where(flipped, w["front"], w["back"])
    args = dict(gap_opening=gap_opening, gap_extension=gap_extension, encoding=encoding)
    a1 = a2 = None
    if len(subseq1["starts"]) or len(subseq1["ends"]):
        a1 = _align_and_extract(adaptor1, actual_starts, subseq_starts=subseq1["starts"], subseq_ends=subseq1["ends"], **args)
    if len(subseq2["starts"]) or len(subseq2["ends"]):
        a2 = _align_and_extract(adaptor2, actual_ends, subseq_starts=subseq2["starts"], subseq_ends=subseq2["ends"], **args)
    return a1, a2


def extractSubseq(aligned, subseq1=None, subseq2=None, number=1e5):
    """R/extractSubseq.R:5-87: arbitrary adaptor sub-ranges (dicts with 1-based "starts"/"ends"), obtained by
    re-aligning and checked against the stored scores -- which is why the device path has to be deterministic and
    independent of chunking / device count.  The reference tests the scores with all.equal; here they are identical."""
    if subseq1 is None and subseq2 is None:
        raise ValueError("at least one of 'subseq1' and 'subseq2' should be specified")
    empty = {"starts": np.zeros(0, np.int32), "ends": np.zeros(0, np.int32)}
    do1, do2 = subseq1 is not None, subseq2 is not None
    subseq1 = subseq1 if do1 else empty
    subseq2 = subseq2 if do2 else empty
    go = aligned["adaptor1"].metadata["gapOpening"]
    ge = aligned["adaptor1"].metadata["gapExtension"]
    adaptor1 = aligned["adaptor1"].metadata["sequence"]
    adaptor2 = aligned["adaptor2"].metadata["sequence"]
    tolerance = aligned.metadata["tolerance"]
    enc = _create_encoding_vector(_qual2class(aligned.metadata["qual.type"]))
    wanted = {nm: k for k, nm in enumerate(aligned.rownames)}
    all1, all2 = [], []
    for reads, _ in _stream(aligned.metadata["filepath"], number):
        keep = np.array([nm in wanted for nm in reads.names], dtype=bool)
        reads = reads[np.nonzero(keep)[0]]
        if len(reads) == 0:
            continue
        m = np.array([wanted[nm] for nm in reads.names], dtype=np.int64)
        a1, a2 = _extract_internal(reads, aligned["reversed"][m], adaptor1, adaptor2, tolerance, subseq1, subseq2, go, ge, enc)
        if do1:
            if a1 is not None and not np.allclose(a1["score"], aligned["adaptor1"]["score"][m], rtol=1.5e-8, atol=0):
                raise RuntimeError("score mismatch from 'aligned' for adaptor 1")
            all1.append(a1["subseq"] if a1 is not None else Frame(nrows=len(reads)))
        if do2:
            if a2 is not None and not np.allclose(a2["score"], aligned["adaptor2"]["score"][m], rtol=1.5e-8, atol=0):
                raise RuntimeError("score mismatch from 'aligned' for adaptor 2")
            all2.append(a2["subseq"] if a2 is not None else Frame(nrows=len(reads)))
    out = {}
    if do1:
        out["adaptor1"] = Frame.rbind(all1) if all1 else Frame(nrows=0)
    if do2:
        out["adaptor2"] = Frame.rbind(all2) if all2 else Frame(nrows=0)
    return out


def qualityAlign(sequences, reference, gapOpening=5, gapExtension=1, edit_only=False, qual_type="phred"):
    """R/qualityAlign.R: global quality-aware alignment of every sequence to one reference (cxx_general_align)."""
    enc = _create_encoding_vector(_qual2class(qual_type))
    ref = str(reference).upper()
    out = native.general_align(sequences, enc, gapOpening, gapExtension, ref, edit_only)
    f = Frame({"score": out[0], "edit": out[1]},
              metadata={"gapOpening": gapOpening, "gapExtension": gapExtension, "reference": reference})
    if not edit_only:
        f["reference"] = np.array(out[2], dtype=object)
        f["query"] = np.array(out[3], dtype=object)
    return f


# --------------------------------------------------------------------------------------------------
# umiGroup (R/umiGroup.R) -- SURVEY 8f-4
# --------------------------------------------------------------------------------------------------
def qualityMask(seq, max_err=None, qual_type="phred"):
    """R/qualityMask.R:5-15 + mask_bad_bases (src/mask_bad_bases.cpp:10-50): bases whose error probability exceeds
    max_err become 'N'.  max_err None (R's NA) or a ReadSet without qualities: sequences unchanged.  Returns a
    quality-less ReadSet (the R function returns a DNAStringSet / character vector)."""
    if not isinstance(seq, ReadSet):
        seq = ReadSet.from_strings(list(seq))
    if max_err is None or not seq.has_quality:
        return ReadSet(seq.seq_pool, seq.seq_off, names=seq.names)
    names, err = _create_encoding_vector(_qual2class(qual_type))
    if not np.array_equal(seq.seq_off, seq.qual_off):
        raise native.SarlaccError("sequence and quality strings should have the same length")
    offset = ord(names[0])
    q = seq.qual_pool[:seq.qual_off[-1]].astype(np.int64) - offset
    if (q < 0).any():
        raise native.SarlaccError("quality cannot be lower than smallest encoded value")
    q = np.minimum(q, len(err) - 1)                       # quality_encoding::to_error clamps like precomputed_cost
    pool = seq.seq_pool[:seq.seq_off[-1]].copy()
    pool[np.asarray(err)[q] > max_err] = ord("N")
    return ReadSet(pool, seq.seq_off, names=seq.names)


def umiGroup(UMI1, threshold1=3, UMI2=None, threshold2=None, max_err=None, groups=None, device=0):
    """R/umiGroup.R:2-23.  groups: None, a vector of group labels (split() semantics: groups ordered by sorted label,
    members in input order) or a list of 1-based index vectors.  Returns a list of int32 arrays (1-based indices)."""
    if threshold2 is None:
        threshold2 = threshold1
    u1 = qualityMask(UMI1, max_err)
    u2 = qualityMask(UMI2, max_err) if UMI2 is not None else None
    n = len(u1)
    if groups is None:
        by_group = [np.arange(1, n + 1, dtype=np.int32)]
    elif isinstance(groups, (list, tuple)) and (len(groups) == 0 or isinstance(groups[0], (list, tuple, np.ndarray))):
        by_group = [np.asarray(g, np.int32) for g in groups]
    else:
        labels = np.asarray(groups)
        uniq, inv = np.unique(labels, return_inverse=True)
        order = np.argsort(inv, kind="stable")
        bounds = np.searchsorted(inv[order], np.arange(len(uniq) + 1))
        by_group = [(order[bounds[k]:bounds[k + 1]] + 1).astype(np.int32) for k in range(len(uniq))]
    return native.umi_group(u1, threshold1, u2, threshold2, by_group, device=device)
