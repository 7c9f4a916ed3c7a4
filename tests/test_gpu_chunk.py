"""GPU: device-resident chunks (sarlacc_chunk_*) -- the device read generator against its host mirror, the in-place
reload, .align_AA_internal / .align_AT_internal on a chunk against the four reference calls composed on the oracle, and
the device threshold selection against the R expression."""
import numpy as np
import pytest

from conftest import VIGNETTE_A1, VIGNETTE_A2

pytestmark = pytest.mark.gpu

S1, E1 = [16, 42], [28, 46]


@pytest.fixture(scope="module", params=["port", "ref"])
def oracle(request):
    return request.getfixturevalue(request.param)


def packed(rs, enc, tol=0, back=0):
    """Host packer (no device) on a ReadSet: the rows every entry point uploads."""
    import ctypes as C
    from sarlacc_b200 import _lib
    ra = _lib.ReadsArg(rs.seq_pool, rs.seq_off, rs.qual_pool, rs.qual_off, 0, False)
    ea = _lib.EncodingArg(*enc)
    stride = 256
    rows = np.zeros((len(rs), stride), np.uint16)
    lens = np.zeros(len(rs), np.int32)
    _lib.check(_lib.lib.sarlacc_pack_rows(ra.ref(), ea.ref(), C.c_int(tol), C.c_int(back), C.c_int(stride), _lib._ptr(rows), _lib._ptr(lens), C.c_int(0)))
    return rows, lens


@pytest.mark.parametrize("barcodes", [None, ["ACGTACGTACGT", "TTTTGGGGCCCC", "GATTACAGATTA"]])
def test_device_generator_matches_host_mirror(enc, barcodes):
    """sarlacc_chunk_load_mock == synth.mock_windows bit for bit: windows, qualities, read widths, strand flips -- and a
    shard generated from another first index tiles the same data set."""
    from sarlacc_b200 import native, synth
    n, first = 6000, 123456789012
    ch = native.Chunk(8192, 250, enc)
    ch.load_mock(n, VIGNETTE_A1, VIGNETTE_A2, seed=2000, first_index=first, barcodes=barcodes)
    rf, lf, width, flips = ch.rows(0)
    rb, lb, _, _ = ch.rows(1)
    front, back, w_host, f_host = synth.mock_windows(n, VIGNETTE_A1, VIGNETTE_A2, seed=2000, first_index=first, barcodes=barcodes)
    pf, plf = packed(front, enc)
    pb, plb = packed(back, enc)
    assert np.array_equal(lf, plf) and np.array_equal(lb, plb) and np.all(lf == 250)
    assert np.array_equal(rf, pf) and np.array_equal(rb, pb)
    assert np.array_equal(width, w_host) and np.array_equal(flips, f_host)
    assert 0.45 < flips.mean() < 0.55 and 5060 < width.mean() < 5120
    # in-place reload with a shifted range
    ch.load_mock(1000, VIGNETTE_A1, VIGNETTE_A2, seed=2000, first_index=first + 500, barcodes=barcodes)
    r2, _, w2, _ = ch.rows(0)
    assert np.array_equal(r2, rf[500:1500]) and np.array_equal(w2, width[500:1500])
    # the decoded windows are the mirror's strings
    dec = synth.unpack_rows(rf, lf)
    assert np.array_equal(dec.seq_pool, front.seq_pool) and np.array_equal(dec.qual_pool, front.qual_pool)
    ch.close()


def test_generated_reads_look_like_mockreads(enc):
    """Error rates of the recipe (R/mockReads.R:73-82): ~3.75 % visible substitutions, ~1 % indels, qualities >= Q12
    with P(Q >= k) falling by 10^-0.1 per step."""
    from sarlacc_b200 import synth
    front, back, widths, flips = synth.mock_windows_device(20000, "ACGT" * 30, "TTGGCCAA" * 4, seed=9, insert_len=1000)
    f = front.seq_pool.reshape(20000, 250)
    want = np.frombuffer(("ACGT" * 30).encode(), np.uint8)
    unflipped = ~flips
    mism = (f[unflipped, :4] != want[None, :4]).mean()            # the first bases: hardly any indel upstream yet
    assert 0.025 < mism < 0.07
    q = front.qual_pool.astype(int) - 33
    assert q.min() == 12 and q.max() <= 93
    tail = [(q >= k).mean() for k in (13, 23, 33)]
    assert abs(tail[1] / tail[0] - 0.1) < 0.02 and abs(tail[2] / tail[1] - 0.1) < 0.03
    assert abs(widths.mean() - (120 + 1000 + 32) * (1 + 0.01 * 1.8)) < 2       # an indel changes the length by (-1+1+2+3+4)/5 on average


def compose(oracle, enc, front, back, widths, go, ge, a1, a2, sec1, sec2):
    """.align_AA_internal + adaptor2 flip from four oracle calls (R/adaptorAlign.R:178-199,66-71)."""
    fa = (front.seq_pool, front.seq_off), (front.qual_pool, front.qual_off)
    ba = (back.seq_pool, back.seq_off), (back.qual_pool, back.qual_off)
    a = oracle.adaptor_align(*fa, enc, go, ge, a1, *sec1)
    b = oracle.adaptor_align(*ba, enc, go, ge, a2, *sec2)
    c = oracle.adaptor_align(*ba, enc, go, ge, a1, *sec1)
    d = oracle.adaptor_align(*fa, enc, go, ge, a2, *sec2)
    rev = (np.maximum(a[0], 0) + np.maximum(b[0], 0)) < (np.maximum(c[0], 0) + np.maximum(d[0], 0))
    r1 = [np.where(rev, c[k], a[k]) for k in range(3)] + [[np.where(rev, c[3][s], a[3][s]) for s in range(len(sec1[0]))],
                                                          [np.where(rev, c[4][s], a[4][s]) for s in range(len(sec1[0]))]]
    r2 = [np.where(rev, d[0], b[0]), widths - np.where(rev, d[1], b[1]) + 1, widths - np.where(rev, d[2], b[2]) + 1,
          [np.where(rev, d[3][s], b[3][s]) for s in range(len(sec2[0]))], [np.where(rev, d[4][s], b[4][s]) for s in range(len(sec2[0]))]]
    return rev, r1, r2


def check_pair(got, exp):
    (w, rev, r1, r2), (erev, e1, e2) = got, exp
    assert np.array_equal(rev, erev)
    for g, e in ((r1, e1), (r2, e2)):
        for k in range(3):
            assert np.array_equal(g[k], e[k]), k
        for s in range(len(e[3])):
            assert np.array_equal(g[3][s], e[3][s]) and np.array_equal(g[4][s], e[4][s])


@pytest.mark.parametrize("go,ge", [(5, 1), (4, 2), (0, 0)])
def test_chunk_adaptor_align_matches_the_four_calls(oracle, enc, go, ge, monkeypatch):
    """sarlacc_chunk_adaptor_align on generated reads == four adaptor_align calls of the oracle + .resolve_strand +
    row selection + adaptor2 flip; several sub-ranges, both scratch parities, reload in place."""
    from sarlacc_b200 import native, synth
    n = 2500
    ch = native.Chunk(4096, 250, enc)
    monkeypatch.setenv("SARLACC_CHUNK", "601")
    monkeypatch.setenv("SARLACC_SPEC_MIN", "512")       # speculative records on these small sub-ranges too
    for first in (0, 77777):
        ch.load_mock(n, VIGNETTE_A1, VIGNETTE_A2, seed=5000, first_index=first)
        front, back, widths, _ = synth.mock_windows(n, VIGNETTE_A1, VIGNETTE_A2, seed=5000, first_index=first)
        got = ch.adaptor_align(go, ge, VIGNETTE_A1, VIGNETTE_A2, (S1, E1), ((), ()))
        assert np.array_equal(got[0], widths)
        check_pair(got, compose(oracle, enc, front, back, widths, go, ge, VIGNETTE_A1, VIGNETTE_A2, (S1, E1), ([], [])))
        assert 0.3 < got[1].mean() < 0.7
    assert "C=18" in ch.last_kernel(0) and "C=22" in ch.last_kernel(1)
    ch.close()


def test_chunk_host_reads_ragged_and_sections_on_both_adaptors(oracle, enc):
    """Host reads through sarlacc_chunk_load_reads: whole reads cut by the packer (ragged lengths, reads shorter than the
    tolerance, empty reads), sections on both adaptors, a generic-kernel adaptor (negative gap opening)."""
    from sarlacc_b200 import native, ReadSet
    from conftest import random_windows
    from oracle import r_level as R
    rng = np.random.default_rng(20261017)
    a1, a2 = "ACGTNNNNACGGTCARRTTGACA", "GGNNTTCCAAGGTT"
    seqs, quals = random_windows(rng, 700, a1, minlen=0, maxlen=400, qlo=2, qhi=40)
    rs = ReadSet.from_strings(seqs, quals)
    tol = 120
    ch = native.Chunk(1024, tol, enc)
    ch.load_reads(rs, tol)
    sec1, sec2 = ([4, 14], [8, 16]), ([2], [4])
    for go, ge in ((5, 1), (-1, 2)):
        got = ch.adaptor_align(go, ge, a1, a2, sec1, sec2)
        w = rs.width().astype(np.int64)
        fw = [s[:tol] for s in seqs]
        fq = [q[:tol] for q in quals]
        bw = [R.revcomp(s[max(0, len(s) - tol):]) for s in seqs]
        bq = [q[max(0, len(q) - tol):][::-1] for q in quals]
        front, back = ReadSet.from_strings(fw, fq), ReadSet.from_strings(bw, bq)
        assert np.array_equal(got[0], w)
        check_pair(got, compose(oracle, enc, front, back, w, go, ge, a1, a2, (list(sec1[0]), list(sec1[1])), (list(sec2[0]), list(sec2[1]))))
    ch.close()


def test_chunk_errors(enc):
    from sarlacc_b200 import native, ReadSet, SarlaccError
    ch = native.Chunk(64, 50, enc)
    with pytest.raises(SarlaccError, match="the chunk holds no reads"):
        ch.adaptor_align(5, 1, "ACGT", "ACGT")
    with pytest.raises(SarlaccError, match="sequence and quality strings should have the same length"):
        ch.load_reads(ReadSet.from_strings(["ACGT", "ACG"], ["5555", "5555"]), 50)
    with pytest.raises(SarlaccError, match="quality cannot be lower than smallest encoded value"):
        ch.load_reads(ReadSet.from_strings(["ACGT", "ACGT"], ["5555", "55 5"]), 50)
    ch.load_reads(ReadSet.from_strings(["ACGTACGTAA"] * 3, ["5" * 10] * 3), 50)
    with pytest.raises(SarlaccError, match="unrecognized base in reference sequence"):
        ch.adaptor_align(5, 1, "ACXT", "ACGT")
    with pytest.raises(SarlaccError, match="chunk runs need two non-empty adaptors"):
        ch.adaptor_align(5, 1, "", "ACGT")
    with pytest.raises(SarlaccError, match="more reads than the chunk's capacity"):
        ch.load_mock(65, "ACGT", "ACGT")
    ch.close()


def test_chunk_scrambled_scores_match_host_scramble(oracle, enc):
    """.align_AT_internal on a chunk: device Fisher-Yates scramble == api._scramble_by_index (rows compared), and the
    kept scores == the oracle's four score-only calls on the host-scrambled windows + .resolve_strand."""
    from sarlacc_b200 import native, synth, api
    n, first = 3000, 4242
    ch = native.Chunk(4096, 250, enc)
    ch.load_mock(n, VIGNETTE_A1, VIGNETTE_A2, seed=5000, first_index=first)
    s1, s2 = ch.scrambled_scores(5, 1, VIGNETTE_A1, VIGNETTE_A2, seed=11, first_index=first)
    front, back, _, _ = synth.mock_windows(n, VIGNETTE_A1, VIGNETTE_A2, seed=5000, first_index=first)
    idx = np.arange(first, first + n, dtype=np.uint64)
    sf, sb = api._scramble_by_index(front, 11, idx, 0), api._scramble_by_index(back, 11, idx, 1)
    assert np.array_equal(ch.rows(2)[0], packed(sf, enc)[0]) and np.array_equal(ch.rows(3)[0], packed(sb, enc)[0])
    assert sorted(sf.seq_strings()[5]) == sorted(front.seq_strings()[5]) and sf.seq_strings()[5] != front.seq_strings()[5]
    fa = (sf.seq_pool, sf.seq_off), (sf.qual_pool, sf.qual_off)
    ba = (sb.seq_pool, sb.seq_off), (sb.qual_pool, sb.qual_off)
    S = oracle.align_score_only(*fa, enc, 5, 1, VIGNETTE_A1)
    E = oracle.align_score_only(*ba, enc, 5, 1, VIGNETTE_A2)
    RS = oracle.align_score_only(*ba, enc, 5, 1, VIGNETTE_A1)
    RE = oracle.align_score_only(*fa, enc, 5, 1, VIGNETTE_A2)
    rev = (np.maximum(S, 0) + np.maximum(E, 0)) < (np.maximum(RS, 0) + np.maximum(RE, 0))
    assert np.array_equal(s1, np.where(rev, RS, S)) and np.array_equal(s2, np.where(rev, RE, E))
    # explicit read indices give the same permutation; unscrambled scores are .get_alignment_scores + .resolve_strand
    t1, t2 = ch.scrambled_scores(5, 1, VIGNETTE_A1, VIGNETTE_A2, seed=11, read_index=idx)
    assert np.array_equal(t1, s1) and np.array_equal(t2, s2)
    u1, u2 = ch.scrambled_scores(5, 1, VIGNETTE_A1, VIGNETTE_A2, scramble=False)
    w, rev_real, r1, r2 = ch.adaptor_align(5, 1, VIGNETTE_A1, VIGNETTE_A2, (S1, E1), ((), ()))
    assert np.array_equal(u1, r1[0]) and np.array_equal(u2, r2[0])
    assert np.median(r1[0]) > np.max(s1) * 0.5 and np.mean(r1[0]) > np.mean(s1) + 20
    ch.close()


def test_chunk_outputs_into_a_result_table_and_device_memory(enc):
    """Chunks of one run write into the columns of one big table (out_pitch) without synchronising in between, into
    page-locked host memory and into device memory alike; the result does not depend on the chunking."""
    import torch
    from sarlacc_b200 import native
    N, cap = 5000, 2048
    ch = native.Chunk(cap, 250, enc)
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()   # noqa: E731
    out = {"reversed": pin(N, torch.uint8), "width": pin(N, torch.int32), "score1": pin(N, torch.float64), "start1": pin(N, torch.int32),
           "end1": pin(N, torch.int32), "sec_start1": pin((2, N), torch.int32), "sec_width1": pin((2, N), torch.int32),
           "start2": pin(N, torch.int32), "end2": pin(N, torch.int32)}
    score2_dev = torch.empty(N, dtype=torch.float64, device="cuda")
    scr1_dev = torch.empty(N, dtype=torch.float64, device="cuda")
    scr2 = pin(N, torch.float64)
    for b0 in range(0, N, cap):
        m = min(cap, N - b0)
        ch.load_mock(m, VIGNETTE_A1, VIGNETTE_A2, seed=5000, first_index=b0)
        o = {k: (v[:, b0:] if v.ndim == 2 else v[b0:]) for k, v in out.items()}
        o["score2"] = score2_dev.data_ptr() + 8 * b0
        ch.adaptor_align(5, 1, VIGNETTE_A1, VIGNETTE_A2, (S1, E1), ((), ()), out=o, out_pitch=N)
        ch.scrambled_scores(5, 1, VIGNETTE_A1, VIGNETTE_A2, seed=3, first_index=b0, score1=scr1_dev.data_ptr() + 8 * b0, score2=scr2[b0:])
    ch.sync()
    big = native.Chunk(N, 250, enc)
    big.load_mock(N, VIGNETTE_A1, VIGNETTE_A2, seed=5000, first_index=0)
    w, rev, r1, r2 = big.adaptor_align(5, 1, VIGNETTE_A1, VIGNETTE_A2, (S1, E1), ((), ()))
    t1, t2 = big.scrambled_scores(5, 1, VIGNETTE_A1, VIGNETTE_A2, seed=3, first_index=0)
    assert np.array_equal(out["width"], w) and np.array_equal(out["reversed"].view(bool), rev)
    assert np.array_equal(out["score1"], r1[0]) and np.array_equal(out["start1"], r1[1]) and np.array_equal(out["end1"], r1[2])
    for s in range(2):
        assert np.array_equal(out["sec_start1"][s], r1[3][s]) and np.array_equal(out["sec_width1"][s], r1[4][s])
    assert np.array_equal(score2_dev.cpu().numpy(), r2[0]) and np.array_equal(out["start2"], r2[1]) and np.array_equal(out["end2"], r2[2])
    assert np.array_equal(scr1_dev.cpu().numpy(), t1) and np.array_equal(scr2, t2)
    # thresholds: device selection == the R expression (host mirror), from host and from device vectors
    from sarlacc_b200 import api
    for err in (0.01, 0.2, 1e-9):
        want = api._compute_threshold(r1[0], t1, err)
        got_h = native.compute_threshold(r1[0], t1, err)
        got_d = native.compute_threshold((torch.from_numpy(r1[0]).cuda().data_ptr(), N), (scr1_dev.data_ptr(), N), err)
        assert (np.isnan(want) and np.isnan(got_h) and np.isnan(got_d)) or (want == got_h == got_d), (err, want, got_h, got_d)
    ch.close()
    big.close()


def test_compute_threshold_edge_cases():
    from sarlacc_b200 import native, api
    rng = np.random.default_rng(5)
    for nr, ns in ((1, 1), (5, 0), (1000, 1000), (100000, 50000)):
        real = np.round(rng.normal(30, 10, nr), 1)          # ties
        scr = np.round(rng.normal(5, 5, ns), 1)
        for err in (0.0, 0.01, 0.5):
            want, got = api._compute_threshold(real, scr, err), native.compute_threshold(real, scr, err)
            assert (np.isnan(want) and np.isnan(got)) or want == got, (nr, ns, err, want, got)
    assert np.isnan(native.compute_threshold(np.zeros(0), np.zeros(3), 0.01))


@pytest.mark.parametrize("mode", ["1", "2"])
def test_speculative_records_do_not_change_results(enc, port, mode, monkeypatch):
    """The strand predictor only decides where traceback records are written (kernels.h: StrandLists).  With every
    prediction inverted (all reads go through the re-run) or every read called unsure (records on both strands) the
    chunk's results are those of the default run and of the four reference calls."""
    from sarlacc_b200 import native, synth
    n = 3000
    ch = native.Chunk(4096, 250, enc)
    ch.load_mock(n, VIGNETTE_A1, VIGNETTE_A2, seed=31, first_index=10 ** 9)
    monkeypatch.setenv("SARLACC_SPEC_MIN", "512")
    base = ch.adaptor_align(5, 1, VIGNETTE_A1, VIGNETTE_A2, (S1, E1), ((), ()))
    monkeypatch.setenv("SARLACC_SPEC_TEST", mode)
    got = ch.adaptor_align(5, 1, VIGNETTE_A1, VIGNETTE_A2, (S1, E1), ((), ()))
    monkeypatch.delenv("SARLACC_SPEC_TEST")
    check_pair(got, (base[1], base[2], base[3]))
    front, back, widths, _ = synth.mock_windows(n, VIGNETTE_A1, VIGNETTE_A2, seed=31, first_index=10 ** 9)
    check_pair(got, compose(port, enc, front, back, widths, 5, 1, VIGNETTE_A1, VIGNETTE_A2, (S1, E1), ([], [])))
    ch.close()
