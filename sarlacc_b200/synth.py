"""mockReads-style synthetic reads (R/mockReads.R), generated from counter-based streams so that every read depends
only on (seed, global read index) -- the same data whatever the chunking, sharding or GPU count.

Molecule = adaptor1 (first N-run <- barcode, other ambiguous positions <- random bases) + uniform ACGT insert +
revcomp(adaptor2) (R/mockReads.R:58-64); per read each base is substituted w.p. sub_rate by a uniform base (:73-74), then
w.p. indel_rate replaced by 0 or 2..max_insert copies of itself (:77-79); qualities are iid Phred of
U(0, sub_rate+indel_rate) (:82); half of the reads are reverse-complemented (:91-92).

Only `tolerance` bases from either end are ever aligned (R/adaptorAlign.R:86-95), so `mock_windows` materialises just
those two windows per read (plus the read width).  It is the HOST MIRROR of the device generator behind
sarlacc_chunk_load_mock (csrc/kernels.cu: mock_windows_kernel): the same 32-bit hash streams evaluated with numpy, bit
for bit (tests/test_gpu_chunk.py).  `mock_reads` materialises whole reads and is meant for small cases (configs[0]).
"""
import numpy as np

from .reads import ReadSet, _COMP

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_U32 = np.uint32


def _hash32(x):
    x = np.asarray(x, dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        x ^= x >> _U32(16)
        x *= _U32(0x21F0AAAD)
        x ^= x >> _U32(15)
        x *= _U32(0x735A2D97)
        x ^= x >> _U32(15)
    return x


def _stream_key(seed, index, field):
    """csrc/kernels.cu: stream_key."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    index = np.asarray(index, dtype=np.uint64)
    k = _hash32(_U32((seed & 0xFFFFFFFF) ^ 0x243F6A88))
    k = _hash32(k ^ _U32(seed >> 32))
    k = _hash32(k ^ (index & np.uint64(0xFFFFFFFF)).astype(np.uint32))
    k = _hash32(k ^ (index >> np.uint64(32)).astype(np.uint32))
    return _hash32(k ^ _U32((int(field) * 0x85EBCA6B + 0xC2B2AE35) & 0xFFFFFFFF))


def _stream_word(key, p):
    """csrc/kernels.cu: stream_word (broadcasts key against p)."""
    with np.errstate(over="ignore"):
        return _hash32(np.asarray(key, np.uint32) ^ (np.asarray(p, np.uint32) * _U32(0x9E3779B1) + _U32(0x7F4A7C15)))


def quality_thresholds(max_err):
    """quality >= k  <=>  word < thr[k] (k = 0..94) and the smallest quality that can occur -- the table
    sarlacc_chunk_load_mock builds (csrc/api.cpp), same expression."""
    thr = np.zeros(95, dtype=np.uint64)
    qmin = 0
    for k in range(95):
        t = 4294967296.0 if k == 0 else float(np.floor(4294967296.0 * (10.0 ** (-(float(k) - 0.5) / 10.0)) / max_err))
        if k == 94:
            t = 0.0
        thr[k] = 0xFFFFFFFF if t >= 4294967295.0 else int(t)
        if k <= 93 and t >= 4294967296.0:
            qmin = k
    return thr, qmin


def _n_runs(adaptor):
    import re
    return [(m.start(), m.end()) for m in re.finditer("N+", adaptor)]


def _mock_window_block(seed, rid, mol, adaptor, run0, barcodes, W, sub_thr, indel_thr, max_insert, thr, qmin):
    """Bases (ASCII) and qualities (Phred+33) of the first W bases of the mutated molecule `mol` of reads rid."""
    n = len(rid)
    P = W + 96                                   # molecule prefix that survives deletions (checked below)
    kfill = _stream_key(seed, rid, 0 + 4 * mol)[:, None]
    kmut = _stream_key(seed, rid, 1 + 4 * mol)[:, None]
    kcnt = _stream_key(seed, rid, 2 + 4 * mol)[:, None]
    kq = _stream_key(seed, rid, 3 + 4 * mol)[:, None]
    pos = np.arange(P, dtype=np.uint32)[None, :]
    base = (_stream_word(kfill, pos) & _U32(3)).astype(np.uint8)
    ad = np.frombuffer(adaptor.encode(), dtype=np.uint8)
    for i, ch in enumerate(ad[:P]):
        code = b"ACGT".find(bytes([ch]))
        if code >= 0:
            base[:, i] = code
    if run0 is not None and mol == 0:
        a, b = run0
        if barcodes:
            bc = np.array([[b"ACGT".index(bytes([c])) for c in x.upper().encode()] for x in barcodes], dtype=np.uint8)
            pick = (_stream_word(kfill[:, 0], 0xFFFFFFFE) % _U32(len(barcodes))).astype(np.int64)
            base[:, a:b] = bc[pick]
        else:
            base[:, a:b] = (_stream_word(kfill[:, 0], 0xFFFFFFFF) & _U32(3)).astype(np.uint8)[:, None]
    u = _stream_word(kmut, pos)
    sub = (u >> _U32(16)) < _U32(sub_thr)
    base = np.where(sub, ((u >> _U32(2)) & _U32(3)).astype(np.uint8), base)
    copies = np.ones((n, P), dtype=np.int64)
    ind = (u & _U32(0xFFFF)) < _U32(indel_thr)
    if ind.any():
        k = (_stream_word(np.broadcast_to(kcnt, (n, P))[ind], np.broadcast_to(pos, (n, P))[ind]) % _U32(max_insert)).astype(np.int64)
        copies[ind] = np.where(k == 0, 0, k + 1)
    tot = copies.sum(axis=1)
    if np.any(tot < W):
        raise ValueError("molecule prefix too short for the requested window (raise P)")
    flat = np.repeat(base.reshape(-1), copies.reshape(-1))
    off = np.zeros(n, dtype=np.int64)
    np.cumsum(tot[:-1], out=off[1:])
    seq = _ACGT[flat[off[:, None] + np.arange(W, dtype=np.int64)[None, :]]]
    h = _stream_word(kq, np.arange(W, dtype=np.uint32)[None, :]).astype(np.uint64)
    asc = thr[1:94][::-1]                          # thresholds of k = 93..1, ascending
    q = np.maximum(len(asc) - np.searchsorted(asc, h, side="right"), qmin)
    return seq, (q + 33).astype(np.uint8)


def width_table(molecule_len, indel_thr):
    """(wlo, wcdf[256]): the number of indels of a whole read is Binomial(molecule_len, indel_thr / 65536), drawn by
    inverting this integer distribution function -- the table sarlacc_chunk_load_mock builds (csrc/api.cpp), same
    operations in the same order on IEEE doubles."""
    pr = float(indel_thr) / 65536.0
    qr = 1.0 - pr
    wlo = max(0, int(float(molecule_len) * pr) - 128)
    x = 1.0
    for _ in range(molecule_len):
        x *= qr
    cdf = 0.0
    out = np.zeros(256, dtype=np.uint64)
    for k in range(wlo + 256):
        cdf += x
        if k >= wlo:
            t = float(np.floor(cdf * 4294967296.0))
            out[k - wlo] = 0xFFFFFFFF if t >= 4294967295.0 else int(t)
        t1 = float(molecule_len - k) * pr
        t2 = float(k + 1) * qr
        x = x * t1
        x = x / t2
    return wlo, out


def _mock_widths(seed, rid, molecule_len, indel_thr, max_insert):
    wlo, wcdf = width_table(molecule_len, indel_thr)
    kw = _stream_key(seed, rid, 9)
    u = _stream_word(kw, 0).astype(np.uint64)
    events = wlo + np.searchsorted(wcdf, u, side="right")        # first k with u < wcdf[k]
    out = np.full(len(rid), molecule_len, dtype=np.int64)
    for e in range(int(events.max()) if len(rid) else 0):
        live = events > e
        k = (_stream_word(kw[live], 1 + e) % _U32(max_insert)).astype(np.int64)
        out[live] += np.where(k == 0, -1, k)
    return out


def mock_windows(n, adaptor1, adaptor2, tolerance=250, seed=2000, insert_len=4908, barcodes=None,
                 sub_rate=0.05, indel_rate=0.01, max_insert=5, first_index=0, block=20000):
    """Front and back windows (as .get_front_and_back would cut them, the back one reverse-complemented) of
    n synthetic reads.  Returns (front ReadSet, back ReadSet, read widths int64[n], flipped bool[n]).
    Read i is a function of (seed, first_index + i) only, so shards generated with different `first_index` tile one big
    data set -- and sarlacc_chunk_load_mock produces the same reads on the device."""
    a1 = adaptor1.upper()
    a2 = adaptor2.upper()
    W = int(tolerance)
    sub_thr, indel_thr = int(np.floor(sub_rate * 65536.0)), int(np.floor(indel_rate * 65536.0))
    thr, qmin = quality_thresholds(sub_rate + indel_rate)
    runs = _n_runs(a1)
    run0 = runs[0] if runs else None
    M = len(a1) + insert_len + len(a2)
    fs, fq, bs, bq, widths, flips = [], [], [], [], [], []
    for b0 in range(0, n, block):
        m = min(block, n - b0)
        rid = np.arange(first_index + b0, first_index + b0 + m, dtype=np.uint64)
        hs, hq = _mock_window_block(seed, rid, 0, a1, run0, barcodes, W, sub_thr, indel_thr, max_insert, thr, qmin)
        ts, tq = _mock_window_block(seed, rid, 1, a2, None, None, W, sub_thr, indel_thr, max_insert, thr, qmin)
        flip = (_stream_word(_stream_key(seed, rid, 8), 0) >> _U32(31)) != 0
        fs.append(np.where(flip[:, None], ts, hs))
        fq.append(np.where(flip[:, None], tq, hq))
        bs.append(np.where(flip[:, None], hs, ts))
        bq.append(np.where(flip[:, None], hq, tq))
        widths.append(_mock_widths(seed, rid, M, indel_thr, max_insert))
        flips.append(flip)

    def build(seqs, quals):
        off = np.arange(n + 1, dtype=np.int64) * W
        if not seqs:
            return ReadSet(np.zeros(0, np.uint8), off, np.zeros(0, np.uint8), off, None)
        return ReadSet(np.concatenate(seqs).reshape(-1), off, np.concatenate(quals).reshape(-1), off, None)

    return (build(fs, fq), build(bs, bq), np.concatenate(widths) if widths else np.zeros(0, np.int64),
            np.concatenate(flips) if flips else np.zeros(0, bool))


def unpack_rows(rows, lens, offset=33):
    """Packed rows (quality index | one-hot base << 8, csrc/kernels.h) back to a ReadSet of ASCII windows (bases that
    were not A, C, G, T read as N).  Inverse of the packer for generated reads."""
    n, stride = rows.shape
    lut = np.full(256, ord("N"), dtype=np.uint8)
    lut[1], lut[2], lut[4], lut[8] = ord("A"), ord("C"), ord("G"), ord("T")
    lens = np.asarray(lens, dtype=np.int64)
    if n and np.all(lens == lens[0]) and lens[0] <= stride:
        w = int(lens[0])
        seq = lut[(rows[:, :w] >> 8).astype(np.uint8)].reshape(-1)
        qual = ((rows[:, :w] & 0xFF) + offset).astype(np.uint8).reshape(-1)
    else:
        keep = np.arange(stride)[None, :] < lens[:, None]
        seq = lut[(rows >> 8).astype(np.uint8)][keep]
        qual = ((rows & 0xFF) + offset).astype(np.uint8)[keep]
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    return ReadSet(seq, off, qual, off, None)


def mock_windows_device(n, adaptor1, adaptor2, tolerance=250, seed=2000, insert_len=4908, barcodes=None,
                        sub_rate=0.05, indel_rate=0.01, max_insert=5, first_index=0, device=0, block=1 << 20):
    """mock_windows produced by the device generator (sarlacc_chunk_load_mock) and copied back: the same reads (the
    parity of the two is a test), at a few million reads per second instead of a few thousand."""
    from . import native
    ch = native.Chunk(min(n, block) or 1, tolerance, native.phred_encoding(), device=device)
    fr, bk, wd, fl = [], [], [], []
    try:
        for b0 in range(0, n, block):
            m = min(block, n - b0)
            ch.load_mock(m, adaptor1, adaptor2, seed=seed, first_index=first_index + b0, insert_len=insert_len, barcodes=barcodes,
                         sub_rate=sub_rate, indel_rate=indel_rate, max_insert=max_insert)
            rf, lf, w, f = ch.rows(0)
            rb, lb, _, _ = ch.rows(1)
            fr.append(unpack_rows(rf, lf))
            bk.append(unpack_rows(rb, lb))
            wd.append(w.astype(np.int64))
            fl.append(f)
    finally:
        ch.close()
    if not fr:
        return mock_windows(0, adaptor1, adaptor2, tolerance=tolerance)
    cat = (lambda parts: parts[0]) if len(fr) == 1 else ReadSet.concat
    return cat(fr), cat(bk), np.concatenate(wd), np.concatenate(fl)


def _fill_adaptor(rng, adaptor, n, barcodes=None):
    """(n, len) uint8 matrix of adaptor copies with the first N-run <- barcode and other N-runs <- random bases."""
    base = np.frombuffer(adaptor.encode(), dtype=np.uint8)
    out = np.tile(base, (n, 1))
    runs = _n_runs(adaptor)
    for k, (a, b) in enumerate(runs):
        if k == 0:
            if barcodes is None:
                pick = rng.integers(0, 4, size=n)
                out[:, a:b] = _ACGT[pick][:, None]          # strrep(nucleotides, barcode.len), :50
            else:
                bc = np.array([np.frombuffer(x.encode(), dtype=np.uint8) for x in barcodes])
                out[:, a:b] = bc[rng.integers(0, len(barcodes), size=n)]
        else:
            out[:, a:b] = _ACGT[rng.integers(0, 4, size=(n, b - a))]
    # any other IUPAC code left in the adaptor is replaced by a random base as well
    other = ~np.isin(out, _ACGT)
    if other.any():
        out[other] = _ACGT[rng.integers(0, 4, size=int(other.sum()))]
    return out


def _phred_quals(rng, shape, max_err):
    p = rng.random(shape) * max_err
    p = np.maximum(p, 1e-12)
    q = np.clip(np.rint(-10.0 * np.log10(p)), 0, 93).astype(np.uint8)
    return (q + 33).astype(np.uint8)


def mock_reads(n, adaptor1, adaptor2, seed=1000, insert_range=(400, 2500), barcodes=None,
               sub_rate=0.05, indel_rate=0.01, max_insert=5, flip_strands=True):
    """Whole synthetic reads (small n).  Returns a ReadSet with names MOLECULE_i:READ_1."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    a1 = adaptor1.upper()
    rc_a2 = _COMP[np.frombuffer(adaptor2.upper().encode(), dtype=np.uint8)][::-1]
    seqs, quals, names = [], [], []
    for i in range(n):
        ins = int(rng.integers(insert_range[0], insert_range[1] + 1))
        mol = np.concatenate([_fill_adaptor(rng, a1, 1, barcodes)[0], _ACGT[rng.integers(0, 4, size=ins)], rc_a2])
        sub = rng.random(len(mol)) < sub_rate
        mol[sub] = _ACGT[rng.integers(0, 4, size=int(sub.sum()))]
        counts = np.ones(len(mol), dtype=np.int64)
        ind = rng.random(len(mol)) < indel_rate
        choices = np.array([0] + list(range(2, max_insert + 1)), dtype=np.int64)
        counts[ind] = choices[rng.integers(0, len(choices), size=int(ind.sum()))]
        read = np.repeat(mol, counts)
        q = _phred_quals(rng, len(read), sub_rate + indel_rate)
        if flip_strands and rng.random() < 0.5:
            read = _COMP[read][::-1]
            q = q[::-1]
        seqs.append(read.tobytes())
        quals.append(q.tobytes())
        names.append("MOLECULE_%d:READ_1" % (i + 1))
    return ReadSet.from_strings(seqs, quals, names)


def mock_barcode_sequences(n, barcodes, seed=3000, sub_rate=0.05, indel_rate=0.01, max_insert=5):
    """Barcode-region subsequences as adaptorAlign would extract them: a mutated copy of a random barcode each."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    bc = np.array([np.frombuffer(x.encode(), dtype=np.uint8) for x in barcodes])
    pick = rng.integers(0, len(barcodes), size=n)
    mol = bc[pick]
    L = mol.shape[1]
    sub = rng.random((n, L)) < sub_rate
    mol[sub] = _ACGT[rng.integers(0, 4, size=int(sub.sum()))]
    counts = np.ones((n, L), dtype=np.int64)
    ind = rng.random((n, L)) < indel_rate
    choices = np.array([0] + list(range(2, max_insert + 1)), dtype=np.int64)
    counts[ind] = choices[rng.integers(0, len(choices), size=int(ind.sum()))]
    flat = np.repeat(mol.reshape(-1), counts.reshape(-1))
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts.sum(axis=1), out=off[1:])
    q = _phred_quals(rng, len(flat), sub_rate + indel_rate)
    return ReadSet(flat, off, q, off, None), pick


def random_barcodes(nb, length=24, min_hamming=8, seed=3000):
    rng = np.random.Generator(np.random.Philox(key=seed + 1))
    out = []
    while len(out) < nb:
        c = _ACGT[rng.integers(0, 4, size=length)]
        if all(int((c != o).sum()) >= min_hamming for o in out):
            out.append(c)
    return [x.tobytes().decode() for x in out]
