"""CPU-only: the host-side mirror of the R drivers (sarlacc_b200/api.py, reads.py, synth.py) against the loop
restatement in oracle/r_level.py and the reference tests' literal answers.  Nothing here launches a kernel."""
import os

import numpy as np
import pytest

from conftest import VIGNETTE_A1, VIGNETTE_A2


def test_setup_subseqs():
    from sarlacc_b200 import api
    # tests/testthat/test-adaptor-align.R:125-127
    for ad, st, en in [("AAAAGGNNNNCCTTTT", [7], [10]), ("AAAAGGYYYYCCTTTT", [7], [10]), ("AAAAGGNNNNCCRRRR", [7, 13], [10, 16]), ("ACGT", [], [])]:
        out = api._setup_subseqs(ad)
        assert out["starts"].tolist() == st and out["ends"].tolist() == en
    out = api._setup_subseqs(VIGNETTE_A1)
    assert out["starts"].tolist() == [17, 43] and out["ends"].tolist() == [28, 46]


def test_front_and_back_windows():
    """tests/testthat/test-adaptor-align.R:130-139 incl. tolerance > read length."""
    from sarlacc_b200 import api, ReadSet
    from oracle import r_level as R
    rng = np.random.default_rng(1)
    seqs = ["".join(rng.choice(list("ACGTN"), size=int(n))) for n in rng.integers(0, 60, size=50)]
    quals = ["".join(chr(33 + int(q)) for q in rng.integers(0, 60, size=len(s))) for s in seqs]
    rs = ReadSet.from_strings(seqs, quals, ["r%d" % i for i in range(len(seqs))])
    for tol in (5, 20, 100):
        w = api._get_front_and_back(rs, tol)
        f, b = R.get_front_and_back(seqs, quals, tol)
        assert w["front"].seq_strings() == [x[0] for x in f] and w["front"].qual_strings() == [x[1] for x in f]
        assert w["back"].seq_strings() == [x[0] for x in b] and w["back"].qual_strings() == [x[1] for x in b]


def test_readset_ops():
    from sarlacc_b200 import ReadSet
    from sarlacc_b200.api import _readset_where
    a = ReadSet.from_strings(["ACGT", "", "GGGTTT"], ["1234", "", "abcdef"], ["x", "y", "z"])
    assert a.width().tolist() == [4, 0, 6]
    assert a.reverse_complement().seq_strings() == ["ACGT", "", "AAACCC"]
    assert a.reverse_complement().qual_strings() == ["4321", "", "fedcba"]
    s = a.subseq(start=[2, 1, 3], width=[2, 0, 4])
    assert s.seq_strings() == ["CG", "", "GTTT"] and s.qual_strings() == ["23", "", "cdef"]
    assert a.subseq(end=[4, 0, 6], width=[1, 0, 2]).seq_strings() == ["T", "", "TT"]
    with pytest.raises(ValueError):
        a.subseq(start=[1, 1, 1], width=[5, 0, 1])
    sub = a[np.array([True, False, True])]
    assert sub.seq_strings() == ["ACGT", "GGGTTT"] and sub.names == ["x", "z"]
    b = ReadSet.from_strings(["TT", "A", "C"], ["!!", "#", "$"])
    w = _readset_where(np.array([True, False, True]), a, b)
    assert w.seq_strings() == ["ACGT", "A", "GGGTTT"] and w.qual_strings() == ["1234", "#", "abcdef"]
    c = ReadSet.concat([a, b])
    assert c.seq_strings() == ["ACGT", "", "GGGTTT", "TT", "A", "C"]
    assert len(ReadSet.empty()) == 0


def test_fastq_round_trip(tmp_path):
    from sarlacc_b200 import ReadSet, read_fastq, write_fastq
    a = ReadSet.from_strings(["ACGT", "GGGTTT", "A"], ["1234", "abcdef", "~"], ["r1", "r2 extra", "r3"])
    p = str(tmp_path / "x.fastq")
    write_fastq(p, a)
    chunks = list(read_fastq(p, number=2))
    assert [len(c) for c in chunks] == [2, 1]
    back = ReadSet.concat(chunks)
    assert back.seq_strings() == a.seq_strings() and back.qual_strings() == a.qual_strings() and back.names == a.names
    open(str(tmp_path / "empty.fastq"), "w").close()
    assert list(read_fastq(str(tmp_path / "empty.fastq"), 10)) == []
    # CRLF line ends, a missing final newline, an empty read, and gzip-compressed input (inflated by the library, zlib)
    with open(str(tmp_path / "crlf.fastq"), "wb") as fh:
        fh.write(b"@a\r\nACGT\r\n+\r\n!!!!\r\n@b\n\n+\n\n@c\nGG\n+c\n##")
    got = ReadSet.concat(list(read_fastq(str(tmp_path / "crlf.fastq"), 2)))
    assert got.seq_strings() == ["ACGT", "", "GG"] and got.qual_strings() == ["!!!!", "", "##"] and got.names == ["a", "b", "c"]
    import gzip
    with gzip.open(str(tmp_path / "x.fastq.gz"), "wb") as fh:
        fh.write(open(p, "rb").read())
    gz = ReadSet.concat(list(read_fastq(str(tmp_path / "x.fastq.gz"), 2)))
    assert gz.seq_strings() == a.seq_strings() and gz.names == a.names
    # the condensed (windows-only) ingest over the same gzip stream: names, widths and both windows of every read
    from sarlacc_b200.reads import read_fastq_condensed
    cond = list(read_fastq_condensed(str(tmp_path / "x.fastq.gz"), 3, 2))
    crs = ReadSet.concat([c[0] for c in cond])
    cw = np.concatenate([c[1] for c in cond])
    assert crs.names == a.names and cw.tolist() == a.width().tolist()
    for got, want in zip(crs.seq_strings(), a.seq_strings()):
        assert got == (want if len(want) <= 6 else want[:3] + want[-3:])
    with open(str(tmp_path / "bad.fastq"), "wb") as fh:
        fh.write(b"@a\nACGT\nIIII\n")
    from sarlacc_b200 import SarlaccError
    with pytest.raises(SarlaccError, match="malformed FASTQ record"):
        list(read_fastq(str(tmp_path / "bad.fastq"), 10))
    with pytest.raises(SarlaccError, match="cannot open FASTQ file"):
        list(read_fastq(str(tmp_path / "missing.fastq"), 10))


def test_resolve_strand_and_thresholds():
    from sarlacc_b200 import api
    from oracle import r_level as R
    rng = np.random.default_rng(3)
    s = [rng.normal(5, 10, size=200) for _ in range(4)]
    s[2][:50] = s[0][:50]
    s[3][:50] = s[1][:50]          # exact ties: fscore == rscore -> not reversed (strict <)
    out = api._resolve_strand(*s)
    rev, final = R.resolve_strand(*s)
    assert out["reversed"].tolist() == rev and np.array_equal(out["scores"], np.array(final))
    assert not out["reversed"][:50].any()
    for trial in range(30):
        real = np.round(rng.normal(30, 10, size=int(rng.integers(1, 80))), 1)
        scr = np.round(rng.normal(10, 8, size=int(rng.integers(1, 80))), 1)
        for err in (0.0, 0.01, 0.2, 1.0):
            a = api._compute_threshold(real, scr, err)
            b = R.compute_threshold(real.tolist(), scr.tolist(), err)
            assert (np.isnan(a) and np.isnan(b)) or a == b
    assert np.isnan(api._compute_threshold([1.0], [5.0], 0.01))


def test_tied_overlap_known_answers():
    """tests/testthat/test-tuning.R:53-59."""
    from sarlacc_b200.api import _tied_overlap
    x = np.arange(1, 11, dtype=float)
    assert _tied_overlap(x, x - 10) == 1
    assert _tied_overlap(x, x) == 0.5
    assert _tied_overlap(x, x - 0.5) == pytest.approx(0.55)
    assert _tied_overlap(x, x + 0.5) == pytest.approx(0.45)
    assert _tied_overlap(x, x + 10) == 0


def test_parallelize_matches_r_rule():
    """R/adaptorAlign.R:126-134: bounds <- seq(1, n, length.out=w+1); ids <- findInterval(seq_len(n), head(bounds,-1))."""
    from sarlacc_b200.api import _parallelize
    for n, w in [(10, 1), (10, 3), (7, 4), (100, 8), (3, 8)]:
        chunks = _parallelize(n, w)
        flat = np.concatenate(chunks)
        assert flat.tolist() == list(range(n))
        bounds = np.linspace(1, n, w + 1)[:-1]
        ids = [int(np.sum(bounds <= i)) for i in range(1, n + 1)]
        assert [len(c) for c in chunks] == [ids.count(k) for k in sorted(set(ids))]
    assert _parallelize(0, 4) == []


def test_scramble_is_a_joint_permutation_and_shard_independent():
    from sarlacc_b200 import api, ReadSet
    rng = np.random.default_rng(5)
    seqs = ["".join(rng.choice(list("ACGT"), size=int(n))) for n in rng.integers(0, 80, size=40)]
    quals = ["".join(chr(40 + (i % 50)) for i in range(len(s))) for s in seqs]
    rs = ReadSet.from_strings(seqs, quals)
    a = api._scramble_input(rs, True, seed=7, first_index=100, stream=0)
    for s, q, s2, q2 in zip(seqs, quals, a.seq_strings(), a.qual_strings()):
        assert sorted(zip(s, q)) == sorted(zip(s2, q2))          # bases and qualities move together
    assert a.seq_strings() != seqs
    b = api._scramble_input(rs, True, seed=7, first_index=100, stream=0)
    assert a.seq_strings() == b.seq_strings()
    # chunk [10:30) scrambled on its own with the right first_index gives the same strings
    part = api._scramble_input(rs[np.arange(10, 30)], True, seed=7, first_index=110, stream=0)
    assert part.seq_strings() == a.seq_strings()[10:30] and part.qual_strings() == a.qual_strings()[10:30]
    assert api._scramble_input(rs, True, seed=8, first_index=100).seq_strings() != a.seq_strings()
    assert api._scramble_by_index(rs, 7, 100 + np.arange(len(rs)), 0).seq_strings() == a.seq_strings()


def test_synthetic_reads_are_shard_independent_and_shaped_like_mockreads():
    from sarlacc_b200 import synth
    f, b, w, fl = synth.mock_windows(3000, VIGNETTE_A1, VIGNETTE_A2, seed=2000, block=1000)
    assert len(f) == 3000 and set(f.width().tolist()) == {250} and set(b.width().tolist()) == {250}
    f2, b2, w2, fl2 = synth.mock_windows(1000, VIGNETTE_A1, VIGNETTE_A2, seed=2000, first_index=1000, block=1000)
    assert f2.seq_strings() == f.seq_strings()[1000:2000] and b2.qual_strings() == b.qual_strings()[1000:2000]
    assert 0.4 < fl.mean() < 0.6
    q = f.qual_pool.astype(int) - 33
    assert q.min() >= 12 and q.max() <= 93                       # -10 log10 U(0, 0.06) >= 12.2
    # un-flipped reads start with adaptor1 (up to 5% substitutions / 1% indels): the constant prefix mostly survives
    s = f.seq_strings()
    hits = sum(1 for i in range(3000) if not fl[i] and s[i].startswith("ACGCAG"))
    assert hits > 0.6 * (~fl).sum()
    assert abs(w.mean() - (70 + 4908 + 22) * 1.018) < 10
    r = synth.mock_reads(20, VIGNETTE_A1, VIGNETTE_A2, seed=1)
    assert len(r) == 20 and r.width().min() > 400


def test_encoding_vector():
    from sarlacc_b200 import api
    names, err = api._create_encoding_vector("PhredQuality")
    assert names[0] == "!" and names[-1] == "~" and len(names) == 94
    assert err[20] == 10.0 ** -2.0 and np.all(np.diff(err) <= 0)
    assert api._qual2class("phred") == "PhredQuality" and api._qual2class("solexa") == "SolexaQuality"
    # Biostrings::encoding(): Illumina '@' (64) .. '~' for Q 0..62; Solexa ';' (59) .. '~' for scores -5..62, error
    # probability 1 / (1 + 10^(q / 10))
    names, err = api._create_encoding_vector("IlluminaQuality")
    assert names[0] == "@" and names[-1] == "~" and len(names) == 63 and err[0] == 1.0 and err[30] == 10.0 ** -3.0
    names, err = api._create_encoding_vector("SolexaQuality")
    assert names[0] == ";" and names[5] == "@" and names[-1] == "~" and len(names) == 68
    assert abs(err[0] - 1.0 / (1.0 + 10.0 ** -0.5)) < 1e-15 and err[5] == 0.5 and abs(err[25] - 1.0 / 101.0) < 1e-15
    assert np.all(np.diff(err) < 0)
    # every table passes the reference's encoding checks (consecutive one-character names, non-increasing errors)
    for cls in ("PhredQuality", "IlluminaQuality", "SolexaQuality"):
        n, e = api._create_encoding_vector(cls)
        assert [ord(b) - ord(a) for a, b in zip(n, n[1:])] == [1] * (len(n) - 1) and np.all(np.diff(e) <= 0)


def _pack(rs_args, enc, tolerance, back, stride, force_scalar, seq_encoding=0):
    import ctypes as C
    from sarlacc_b200 import _lib
    ra = _lib.ReadsArg(*rs_args, seq_encoding, False)
    ea = _lib.EncodingArg(*enc)
    rows = np.full((ra.n, stride), 0xABCD, np.uint16)
    lens = np.zeros(ra.n, np.int32)
    rc = _lib.lib.sarlacc_pack_rows(ra.ref(), ea.ref(), C.c_int(tolerance), C.c_int(back), C.c_int(stride),
                                    _lib._ptr(rows), _lib._ptr(lens), C.c_int(force_scalar))
    return rc, rows, lens


def test_vector_packer_matches_scalar_and_numpy():
    """The AVX2 packer, the table packer and a numpy restatement of DNA_input decode + the quality clamp
    (src/DNA_input.cpp:64-75, src/reference_align.cpp:215-221, R/adaptorAlign.R:86-95) agree on every entry."""
    from sarlacc_b200 import _lib
    from sarlacc_b200.native import phred_encoding
    rng = np.random.default_rng(7)
    n = 300
    lens_true = rng.integers(0, 400, size=n)
    lens_true[:8] = [0, 1, 31, 32, 33, 63, 64, 65]
    alphabet = np.frombuffer(b"ACGTNacgtRYKM-", np.uint8)
    off = np.concatenate([[0], np.cumsum(lens_true)]).astype(np.int64)
    seq = alphabet[rng.integers(0, len(alphabet), size=off[-1])]
    qual = rng.integers(33, 127, size=off[-1]).astype(np.uint8)
    onehot = np.zeros(256, np.uint16)
    for ch, v in zip(b"ACGT", (1, 2, 4, 8)):
        onehot[ch] = v
    comp = np.zeros(256, np.uint16)
    for ch, v in zip(b"ACGT", (8, 4, 2, 1)):
        comp[ch] = v
    for nenc in (94, 42):       # 42: qualities above the table are clamped to its last entry
        enc = phred_encoding(nenc)
        for tol, back in ((0, 0), (250, 0), (250, 1), (40, 1), (31, 1)):
            stride = 400
            rc_v, rows_v, lens_v = _pack((seq, off, qual, off), enc, tol, back, stride, 0)
            rc_s, rows_s, lens_s = _pack((seq, off, qual, off), enc, tol, back, stride, 1)
            assert rc_v == 0 and rc_s == 0
            assert np.array_equal(lens_v, lens_s)
            for i in range(n):
                ln = int(lens_true[i]) if tol == 0 else min(tol, int(lens_true[i]))
                assert lens_v[i] == ln
                s = seq[off[i]:off[i + 1]]
                q = qual[off[i]:off[i + 1]]
                if back:
                    s, q = s[len(s) - ln:][::-1], q[len(q) - ln:][::-1]
                    base = comp[s]
                else:
                    s, q = s[:ln], q[:ln]
                    base = onehot[s]
                want = np.minimum(q.astype(np.int32) - 33, nenc - 1).astype(np.uint16) | (base << 8)
                assert np.array_equal(rows_s[i, :ln], want), (nenc, tol, back, i)
                assert np.array_equal(rows_v[i, :ln], want), (nenc, tol, back, i)
                assert (rows_v[i, ln:] == 0xABCD).all()      # nothing written past the window
    # Biostrings byte codes (A=1, C=2, G=4, T=8; IUPAC = OR; gaps >= 16)
    codes = np.array([1, 2, 4, 8, 15, 3, 5, 16, 32], np.uint8)
    seq_b = codes[rng.integers(0, len(codes), size=off[-1])]
    for back in (0, 1):
        rc_v, rows_v, _ = _pack((seq_b, off, qual, off), phred_encoding(), 100, back, 128, 0, seq_encoding=_lib.SEQ_BIOSTRINGS)
        rc_s, rows_s, lens_s = _pack((seq_b, off, qual, off), phred_encoding(), 100, back, 128, 1, seq_encoding=_lib.SEQ_BIOSTRINGS)
        assert rc_v == 0 and rc_s == 0
        for i in range(n):
            assert np.array_equal(rows_v[i, :lens_s[i]], rows_s[i, :lens_s[i]])
    # a quality below the offset fails with the reference's message on both paths, wherever it sits in the window
    for pos in (0, 5, 31, 32, 100, 249):
        q2 = qual.copy()
        q2[off[20] + pos] = 32      # read 20 is longer than 250? make sure below
        if lens_true[20] <= pos:
            continue
        for fs in (0, 1):
            rc, _, _ = _pack((seq, off, q2, off), phred_encoding(), 0, 0, 400, fs)
            assert rc != 0 and "quality cannot be lower than smallest encoded value" in _lib.last_error()


def test_condensed_fastq_ingest(tmp_path):
    """sarlacc_fastq_next_condensed (parallel, windows only) against the sequential whole-read reader: same names,
    widths and first/last `keep` bases for any thread count and chunk size; quality lines that start with '@' or '+',
    CRLF line ends, blank lines, a missing final newline; malformed records are reported."""
    from sarlacc_b200 import read_fastq, read_fastq_condensed, SarlaccError
    rng = np.random.default_rng(12)
    n = 3000
    recs = []
    for i in range(n):
        ln = int(rng.integers(0, 1500)) if i % 50 else int(rng.integers(0, 12))
        s = "".join(rng.choice(list("ACGTN"), ln))
        q = "".join(chr(int(c)) for c in rng.integers(33, 127, ln))
        if ln and i % 7 == 0:
            q = "@" + q[1:]                       # a quality line that looks like a header
        if ln and i % 11 == 0:
            q = "+" + q[1:]
        recs.append(("read_%d some comment" % i, s, q))
    for crlf, final_nl in ((False, True), (True, True), (False, False)):
        nl = "\r\n" if crlf else "\n"
        text = "".join("@%s%s%s%s+%s%s%s" % (h, nl, s, nl, nl, q, nl) + ("\n" if k % 500 == 499 else "") for k, (h, s, q) in enumerate(recs))
        if not final_nl:
            text = text.rstrip("\r\n")
        path = tmp_path / ("x_%d_%d.fastq" % (crlf, final_nl))
        path.write_bytes(text.encode("latin-1"))
        whole = list(read_fastq(str(path)))
        assert sum(len(c) for c in whole) == n
        names = [x for c in whole for x in c.names]
        seqs = [x for c in whole for x in c.seq_strings()]
        quals = [x for c in whole for x in c.qual_strings()]
        assert names == [h for h, _, _ in recs] and seqs == [s for _, s, _ in recs] and quals == [q for _, _, q in recs]
        for keep, number, threads in ((250, None, 1), (250, None, 5), (40, 777, 3), (5, 64, 8), (250, 1, 2)):
            if number == 1 and (crlf or not final_nl):
                continue
            got_n, got_s, got_q, got_w = [], [], [], []
            for rs, w in read_fastq_condensed(str(path), keep, number, nthreads=threads):
                assert number is None or len(rs) <= number
                got_n += rs.names
                got_s += rs.seq_strings()
                got_q += rs.qual_strings()
                got_w += w.tolist()
            cond = lambda x: x if len(x) <= 2 * keep else x[:keep] + x[-keep:]   # noqa: E731
            assert got_n == names and got_w == [len(s) for s in seqs]
            assert got_s == [cond(s) for s in seqs] and got_q == [cond(q) for q in quals]
    bad = tmp_path / "bad.fastq"
    bad.write_bytes(b"@r1\nACGT\n+\n!!!!\n@r2\nACGT\n+\n!!!\n")
    with pytest.raises(SarlaccError, match="sequence and quality lengths differ"):
        list(read_fastq_condensed(str(bad), 10))
    bad.write_bytes(b"@r1\nACGT\n+\n!!!!\nr2\nACGT\n+\n!!!!\n")
    with pytest.raises(SarlaccError, match="header does not start with '@'"):
        list(read_fastq_condensed(str(bad), 10))
    empty = tmp_path / "empty.fastq"
    empty.write_bytes(b"")
    assert list(read_fastq_condensed(str(empty), 10)) == []


def test_sample_reads_is_a_uniform_sample_in_stream_order():
    """FastqSampler's job (R/tuneAlignment.R:21): `number` distinct reads, every read equally likely, stream order kept,
    the same sample for the same seed whatever the chunking."""
    from sarlacc_b200 import api, ReadSet
    n = 3000
    rs = ReadSet.from_strings(["ACGT" * 3] * n, ["5555" * 3] * n, ["r%d" % i for i in range(n)])
    s = api._sample_reads(rs, 500, 250, seed=1)
    pos = [int(x[1:]) for x in s.names]
    assert len(pos) == 500 and len(set(pos)) == 500 and pos == sorted(pos)
    assert api._sample_reads(rs, 500, 250, seed=1).names == s.names and api._sample_reads(rs, 500, 250, seed=2).names != s.names
    hits = np.zeros(n)
    for seed in range(40):
        for x in api._sample_reads(rs, 300, 250, seed=seed).names:
            hits[int(x[1:])] += 1
    assert abs(hits[:1000].mean() - hits[2000:].mean()) < 0.6 and 3.0 < hits.mean() < 5.0     # 40 * 300 / 3000 = 4 per read
    assert len(api._sample_reads(rs, 5000, 250, seed=0)) == n                                   # fewer reads than asked for: all of them


def test_bases_as_four_bit_codes():
    """sarlacc_pack_bases (the host half of SARLACC_PACK_SEQ=1): every byte value, odd lengths, the vector and the table
    loop, ASCII and Biostrings codes == the packer table (A, C, G, T -> 1, 2, 4, 8; anything else, lower case included,
    -> 0)."""
    import ctypes as C
    from sarlacc_b200 import _lib
    rng = np.random.default_rng(5)
    tables = {0: {65: 1, 67: 2, 71: 4, 84: 8}, 1: {1: 1, 2: 2, 4: 4, 8: 8}}
    for enc_id, tab in tables.items():
        lut = np.zeros(256, np.uint8)
        for k, v in tab.items():
            lut[k] = v
        for n in (0, 1, 2, 63, 64, 65, 127, 4097, 300001):
            seq = rng.integers(0, 256, n).astype(np.uint8)
            if n >= 256:
                seq[:256] = np.arange(256)
                seq[256:] = np.frombuffer(b"ACGTNacgt\x01\x02\x04\x08", np.uint8)[rng.integers(0, 13, n - 256)]
            codes = lut[seq]
            if n & 1:
                codes = np.append(codes, np.uint8(0))
            exp = (codes[0::2] | (codes[1::2] << 4)).astype(np.uint8)
            for scalar in (0, 1):
                out = np.full((n + 1) // 2 + 8, 0xEE, np.uint8)
                _lib.check(_lib.lib.sarlacc_pack_bases(_lib._ptr(seq), C.c_int64(n), C.c_int(enc_id), _lib._ptr(out), C.c_int(scalar)))
                assert np.array_equal(out[:(n + 1) // 2], exp), (enc_id, n, scalar)
                assert np.all(out[(n + 1) // 2:] == 0xEE)
