"""GPU: the host-side mirror of the R drivers (adaptorAlign / getAdaptorThresholds / barcodeAlign / tuneAlignment /
qualityAlign) through the CUDA library, against the loop restatement on the CPU oracle."""
import numpy as np
import pytest

from conftest import VIGNETTE_A1, VIGNETTE_A2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def port(both_oracles):
    """Both CPU oracles in turn: the restatement and the reference's own C++ (tests/conftest.py: both_oracles)."""
    return both_oracles


def compare_adaptor_align(out, exp):
    assert len(out) == len(exp)
    for side in ("adaptor1", "adaptor2"):
        f = out[side]
        assert np.array_equal(f["score"], np.array([e[side]["score"] for e in exp]))
        assert f["start"].tolist() == [e[side]["start"] for e in exp]
        assert f["end"].tolist() == [e[side]["end"] for e in exp]
        nsub = len(exp[0][side]["subseq"]) if exp else 0
        for k in range(nsub):
            sub = f["subseq"]["Sub%d" % (k + 1)]
            assert sub.seq_strings() == [e[side]["subseq"][k][0] for e in exp]
            assert sub.qual_strings() == [e[side]["subseq"][k][1] for e in exp]
    assert out["reversed"].tolist() == [e["reversed"] for e in exp]
    assert out["read.width"].tolist() == [e["read.width"] for e in exp]


def test_adaptor_align_known_answer(port, enc):
    """tests/testthat/test-adaptor-align.R:186-206."""
    from sarlacc_b200 import api, ReadSet
    from oracle import r_level as R
    read = "AACGTAACGTACGTACGTGGGGGGG"
    rs = ReadSet.from_strings([read, R.revcomp(read)], ["5" * 25, "5" * 25], ["fwd", "rev"])
    out = api.adaptorAlign("AANNNAA", "CCCCCCC", rs)
    assert out["reversed"].tolist() == [False, True]
    assert out["adaptor1"]["start"].tolist() == [1, 1] and out["adaptor1"]["end"].tolist() == [7, 7]
    assert out["adaptor2"]["start"].tolist() == [25, 25] and out["adaptor2"]["end"].tolist() == [19, 19]
    assert out["adaptor1"]["subseq"]["Sub1"].seq_strings() == ["CGT", "CGT"]
    assert out["adaptor1"]["score"][0] == out["adaptor1"]["score"][1]
    assert out.rownames == ["fwd", "rev"] and out.metadata["tolerance"] == 250
    assert out["adaptor1"].metadata == {"sequence": "AANNNAA", "gapOpening": 5, "gapExtension": 1}
    # empty input -> 0-row result (:209-211)
    empty = api.adaptorAlign("AANNNAA", "CCCCCCC", ReadSet.empty())
    assert len(empty["read.width"]) == 0 and len(empty["adaptor1"]["score"]) == 0


def test_adaptor_align_mock_reads(port, enc, tmp_path):
    """configs[0] in miniature: mockReads-style whole reads through a FASTQ file, chunked streaming, both
    strands, vignette adaptors with UMI/barcode N-runs."""
    from sarlacc_b200 import api, synth, write_fastq
    from oracle import r_level as R
    reads = synth.mock_reads(240, VIGNETTE_A1, VIGNETTE_A2, seed=1000)
    path = str(tmp_path / "mock.fastq")
    write_fastq(path, reads)
    out = api.adaptorAlign(VIGNETTE_A1, VIGNETTE_A2, path, number=100)
    exp = R.adaptor_align_R(port, enc, VIGNETTE_A1, VIGNETTE_A2, reads.seq_strings(), reads.qual_strings())
    compare_adaptor_align(out, exp)
    compare_adaptor_align(api.adaptorAlign(VIGNETTE_A1, VIGNETTE_A2, path, number=100, fused=False), exp)
    assert out.rownames == reads.names
    assert 0.2 < out["reversed"].mean() < 0.8
    # most reads carry both adaptors: scores well above zero, 12-base barcode and 4-base UMI extracted
    assert np.median(out["adaptor1"]["score"]) > 20
    assert np.median(out["adaptor1"]["subseq"]["Sub2"].width()) == 4
    # other tolerance / penalties, in-memory input
    out2 = api.adaptorAlign(VIGNETTE_A1, VIGNETTE_A2, reads, tolerance=120, gapOpening=4, gapExtension=2, number=77)
    exp2 = R.adaptor_align_R(port, enc, VIGNETTE_A1, VIGNETTE_A2, reads.seq_strings(), reads.qual_strings(), tolerance=120, go=4, ge=2)
    compare_adaptor_align(out2, exp2)


def test_fused_windows_entry(port, enc, monkeypatch):
    """sarlacc_adaptor_align_windows == the four reference calls + .resolve_strand + selection + flip."""
    from sarlacc_b200 import native, synth, SarlaccError
    from oracle import r_level as R
    n = 3000
    front, back, widths, _ = synth.mock_windows(n, VIGNETTE_A1, VIGNETTE_A2, seed=77)
    s1, e1 = [16, 42], [28, 46]
    monkeypatch.setenv("SARLACC_CHUNK", "700")
    monkeypatch.setenv("SARLACC_SPEC_MIN", "512")       # speculative records on these small chunks too
    rev, r1, r2 = native.adaptor_align_windows(front, back, enc, 5, 1, VIGNETTE_A1, VIGNETTE_A2, (s1, e1), ((), ()), read_width=widths)
    a = native.adaptor_align(front, enc, 5, 1, VIGNETTE_A1, s1, e1)
    b = native.adaptor_align(back, enc, 5, 1, VIGNETTE_A2)
    c = native.adaptor_align(back, enc, 5, 1, VIGNETTE_A1, s1, e1)
    d = native.adaptor_align(front, enc, 5, 1, VIGNETTE_A2)
    exp_rev, _ = R.resolve_strand(a[0], b[0], c[0], d[0])
    exp_rev = np.array(exp_rev)
    assert np.array_equal(rev, exp_rev) and 0.3 < rev.mean() < 0.7
    for k in range(3):
        assert np.array_equal(r1[k], np.where(exp_rev, c[k], a[k]))
    for s in range(2):
        assert np.array_equal(r1[3][s], np.where(exp_rev, c[3][s], a[3][s])) and np.array_equal(r1[4][s], np.where(exp_rev, c[4][s], a[4][s]))
    assert np.array_equal(r2[0], np.where(exp_rev, d[0], b[0]))
    assert np.array_equal(r2[1], widths - np.where(exp_rev, d[1], b[1]) + 1) and np.array_equal(r2[2], widths - np.where(exp_rev, d[2], b[2]) + 1)
    # degenerate adaptor and error paths go through the same entry
    rev0, q1, q2 = native.adaptor_align_windows(front[np.arange(50)], back[np.arange(50)], enc, 5, 1, "", VIGNETTE_A2)
    assert np.all(q1[0] == 0) and np.array_equal(q2[0], np.where(rev0, d[0][:50], b[0][:50]))
    with pytest.raises(SarlaccError, match="unrecognized base in reference sequence"):
        native.adaptor_align_windows(front[np.arange(5)], back[np.arange(5)], enc, 5, 1, VIGNETTE_A1, "ACXT")
    with pytest.raises(SarlaccError, match="front and back windows should have the same length"):
        native.adaptor_align_windows(front[np.arange(5)], back[np.arange(4)], enc, 5, 1, VIGNETTE_A1, VIGNETTE_A2)


def test_pair_pipeline_default_chunking(enc):
    """The host-buffer pipeline at a size that needs several default chunks (quarter-size first chunk, whole grid-fulls
    after it, three slots in rotation): sarlacc_adaptor_align_windows over 300 k reads == four resident passes +
    .resolve_strand + selection, and == itself with a different chunking."""
    import os
    from sarlacc_b200 import native, synth
    from oracle import r_level as R
    n = 300000
    front, back, widths, _ = synth.mock_windows_device(n, VIGNETTE_A1, VIGNETTE_A2, seed=31)
    s1, e1 = [16, 42], [28, 46]
    rev, r1, r2 = native.adaptor_align_windows(front, back, enc, 5, 1, VIGNETTE_A1, VIGNETTE_A2, (s1, e1), ((), ()), read_width=widths)
    res = {}
    for key, rs, ad, sec in (("a", front, VIGNETTE_A1, (s1, e1)), ("b", back, VIGNETTE_A2, ((), ())),
                             ("c", back, VIGNETTE_A1, (s1, e1)), ("d", front, VIGNETTE_A2, ((), ()))):
        r = native.Resident(rs, enc)
        r.align(r.MODE_TRACE_LOCAL, 5, 1, ad, *sec)
        res[key] = r.fetch()
        r.close()
    a, b, c, d = res["a"], res["b"], res["c"], res["d"]
    f = np.maximum(a[0], 0) + np.maximum(b[0], 0)
    rr = np.maximum(c[0], 0) + np.maximum(d[0], 0)
    exp_rev = f < rr                                              # R/adaptorAlign.R:112-122
    assert np.array_equal(rev, exp_rev)
    for k in range(3):
        assert np.array_equal(r1[k], np.where(exp_rev, c[k], a[k]))
    for s in range(2):
        assert np.array_equal(r1[3][s], np.where(exp_rev, c[3][s], a[3][s])) and np.array_equal(r1[4][s], np.where(exp_rev, c[4][s], a[4][s]))
    assert np.array_equal(r2[0], np.where(exp_rev, d[0], b[0]))
    assert np.array_equal(r2[1], widths - np.where(exp_rev, d[1], b[1]) + 1) and np.array_equal(r2[2], widths - np.where(exp_rev, d[2], b[2]) + 1)
    os.environ["SARLACC_CHUNK"] = "50001"
    try:
        rev2, q1, q2 = native.adaptor_align_windows(front, back, enc, 5, 1, VIGNETTE_A1, VIGNETTE_A2, (s1, e1), ((), ()), read_width=widths)
    finally:
        del os.environ["SARLACC_CHUNK"]
    assert np.array_equal(rev, rev2)
    for k in range(3):
        assert np.array_equal(r1[k], q1[k]) and np.array_equal(r2[k], q2[k])


def test_pinned_inputs_take_the_zero_copy_path(enc):
    """Pools in pinned host memory are DMA'd straight to the device packer (no staging copy): identical results to
    pageable pools, for the fused entry and the single-adaptor entries, including a quality error found on the device."""
    import torch
    from sarlacc_b200 import native, synth, ReadSet, SarlaccError
    n = 150000
    front, back, widths, _ = synth.mock_windows_device(n, VIGNETTE_A1, VIGNETTE_A2, seed=55)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()   # noqa: E731
    pf = ReadSet(pin(front.seq_pool), front.seq_off, pin(front.qual_pool), front.qual_off, front.names)
    pb = ReadSet(pin(back.seq_pool), back.seq_off, pin(back.qual_pool), back.qual_off, back.names)
    s1, e1 = [16, 42], [28, 46]
    a = native.adaptor_align_windows(front, back, enc, 5, 1, VIGNETTE_A1, VIGNETTE_A2, (s1, e1), ((), ()), read_width=widths)
    b = native.adaptor_align_windows(pf, pb, enc, 5, 1, VIGNETTE_A1, VIGNETTE_A2, (s1, e1), ((), ()), read_width=widths)
    assert np.array_equal(a[0], b[0])
    for x, y in ((a[1], b[1]), (a[2], b[2])):
        for k in range(3):
            assert np.array_equal(x[k], y[k])
        for k in range(len(x[3])):
            assert np.array_equal(x[3][k], y[3][k]) and np.array_equal(x[4][k], y[4][k])
    c = native.adaptor_align(front, enc, 5, 1, VIGNETTE_A1, s1, e1)
    d = native.adaptor_align(pf, enc, 5, 1, VIGNETTE_A1, s1, e1)
    assert all(np.array_equal(c[k], d[k]) for k in range(3))
    assert np.array_equal(native.adaptor_align_score_only(pb, enc, 5, 1, VIGNETTE_A2), native.adaptor_align_score_only(back, enc, 5, 1, VIGNETTE_A2))
    # a quality below the offset somewhere in the middle: reported by the device packer, same message
    pf.qual_pool[pf.qual_off[120000] + 17] = 32
    with pytest.raises(SarlaccError, match="quality cannot be lower than smallest encoded value"):
        native.adaptor_align(pf, enc, 5, 1, VIGNETTE_A1, s1, e1)
    with pytest.raises(SarlaccError, match="quality cannot be lower than smallest encoded value"):
        native.adaptor_align_windows(pf, pb, enc, 5, 1, VIGNETTE_A1, VIGNETTE_A2, (s1, e1), ((), ()), read_width=widths)


def test_fused_whole_read_entry(port, enc):
    """sarlacc_adaptor_align_reads: windows cut and reverse-complemented by the packer == .get_front_and_back + the
    four calls, for reads shorter and longer than the tolerance, odd characters and Biostrings byte codes."""
    from sarlacc_b200 import api, native, synth, ReadSet, SEQ_BIOSTRINGS
    from oracle import r_level as R
    reads = synth.mock_reads(150, VIGNETTE_A1, VIGNETTE_A2, seed=41, insert_range=(10, 900))
    seqs, quals = reads.seq_strings(), reads.qual_strings()
    seqs[3] = seqs[3][:40] + "NNRYK" + seqs[3][45:]          # IUPAC codes in a read: never match, complement irrelevant
    seqs[5], quals[5] = "", ""
    rs = ReadSet.from_strings(seqs, quals, reads.names)
    for tol in (250, 60):
        width, rev, r1, r2 = native.adaptor_align_reads(rs, tol, enc, 5, 1, VIGNETTE_A1, VIGNETTE_A2, ([16, 42], [28, 46]), ((), ()))
        exp = R.adaptor_align_R(port, enc, VIGNETTE_A1, VIGNETTE_A2, seqs, quals, tolerance=tol)
        assert width.tolist() == [len(x) for x in seqs] and rev.tolist() == [e["reversed"] for e in exp]
        assert np.array_equal(r1[0], np.array([e["adaptor1"]["score"] for e in exp]))
        assert r1[1].tolist() == [e["adaptor1"]["start"] for e in exp] and r1[2].tolist() == [e["adaptor1"]["end"] for e in exp]
        assert np.array_equal(r2[0], np.array([e["adaptor2"]["score"] for e in exp]))
        assert r2[1].tolist() == [e["adaptor2"]["start"] for e in exp] and r2[2].tolist() == [e["adaptor2"]["end"] for e in exp]
        out = api.adaptorAlign(VIGNETTE_A1, VIGNETTE_A2, rs, tolerance=tol, number=64)
        compare_adaptor_align(out, exp)
    code = {"A": 1, "C": 2, "G": 4, "T": 8, "N": 15, "R": 5, "Y": 10, "K": 12}
    coded = ReadSet.from_strings([bytes(code[c] for c in s) for s in seqs], quals)
    w2, rev2, q1, q2 = native.adaptor_align_reads(coded, 250, enc, 5, 1, VIGNETTE_A1, VIGNETTE_A2, seq_encoding=SEQ_BIOSTRINGS)
    width, rev, r1, r2 = native.adaptor_align_reads(rs, 250, enc, 5, 1, VIGNETTE_A1, VIGNETTE_A2)
    assert np.array_equal(rev2, rev) and np.array_equal(q1[0], r1[0]) and np.array_equal(q2[1], r2[1])
    with pytest.raises(native.SarlaccError, match="two non-empty adaptors"):
        native.adaptor_align_reads(rs, 250, enc, 5, 1, "", VIGNETTE_A2)


def test_get_adaptor_thresholds(port, enc):
    from sarlacc_b200 import api, synth
    from oracle import r_level as R
    reads = synth.mock_reads(150, VIGNETTE_A1, VIGNETTE_A2, seed=5)
    aligned = api.adaptorAlign(VIGNETTE_A1, VIGNETTE_A2, reads, number=60)
    th = api.getAdaptorThresholds(aligned, error=0.01, number=50, seed=3)
    # identity is defined from the scrambled windows onwards: rebuild them with the same keyed permutation
    w = api._get_front_and_back(reads, 250)
    sf = api._scramble_input(w["front"], True, 3, 0, 0)
    sb = api._scramble_input(w["back"], True, 3, 0, 1)
    t1, t2, s1, s2 = R.adaptor_thresholds_R(port, enc, VIGNETTE_A1, VIGNETTE_A2,
                                            list(zip(sf.seq_strings(), sf.qual_strings())), list(zip(sb.seq_strings(), sb.qual_strings())),
                                            aligned["adaptor1"]["score"].tolist(), aligned["adaptor2"]["score"].tolist(), 5, 1, 0.01)
    assert np.array_equal(th["scores1"]["scrambled"], np.array(s1)) and np.array_equal(th["scores2"]["scrambled"], np.array(s2))
    assert (np.isnan(t1) and np.isnan(th["threshold1"])) or t1 == th["threshold1"]
    assert (np.isnan(t2) and np.isnan(th["threshold2"])) or t2 == th["threshold2"]
    assert np.array_equal(th["scores1"]["reads"], aligned["adaptor1"]["score"])
    # real adaptors score far above scrambled windows
    assert np.median(th["scores1"]["reads"]) > np.max(th["scores1"]["scrambled"])
    # the host (numpy) scramble + four reference-style calls give the same numbers as the device scramble
    th_host = api.getAdaptorThresholds(aligned, error=0.01, number=50, seed=3, device_scramble=False)
    assert np.array_equal(th_host["scores1"]["scrambled"], th["scores1"]["scrambled"])
    assert np.array_equal(th_host["scores2"]["scrambled"], th["scores2"]["scrambled"])
    # chunking must not change anything
    th2 = api.getAdaptorThresholds(aligned, error=0.01, number=1000, seed=3)
    assert np.array_equal(th2["scores1"]["scrambled"], th["scores1"]["scrambled"]) and th2["threshold2"] == th["threshold2"] or np.isnan(th["threshold2"])


def test_device_scramble_matches_host_permutation(enc):
    """sarlacc_resident_scrambled == api._scramble_input row for row (same keyed permutation), ragged lengths included."""
    from sarlacc_b200 import api, native, ReadSet
    from conftest import random_windows
    rng = np.random.default_rng(12)
    seqs, quals = random_windows(rng, 300, VIGNETTE_A2, 0, 260, 0, 60)
    rs = ReadSet.from_strings(seqs, quals)
    r = native.Resident(rs, enc)
    for seed, first, stream in [(0, 0, 0), (7, 1000, 1), (2 ** 40 + 5, 123456789, 0)]:
        d = r.scrambled(seed, first_index=first, stream_id=stream)
        rows, lens = d.rows()
        h = native.Resident(api._scramble_input(rs, True, seed, first, stream), enc)
        hrows, hlens = h.rows()
        assert np.array_equal(lens, hlens)
        for i in range(len(rs)):
            assert np.array_equal(rows[i, :lens[i]], hrows[i, :lens[i]]), (seed, i)
        d.close()
        h.close()
    idx = rng.permutation(5000)[:300].astype(np.uint64)
    d = r.scrambled(9, read_index=idx, stream_id=1)
    h = native.Resident(api._scramble_by_index(rs, 9, idx, 1), enc)
    a, la = d.rows()
    b, lb = h.rows()
    assert all(np.array_equal(a[i, :la[i]], b[i, :lb[i]]) for i in range(len(rs)))
    r.close()


def test_barcode_align(port, enc):
    from sarlacc_b200 import api, synth
    from oracle import r_level as R
    barcodes = synth.random_barcodes(12, 24, 8, seed=3000)
    seqs, pick = synth.mock_barcode_sequences(400, barcodes, seed=3000)
    out = api.barcodeAlign(seqs, barcodes)
    cid, cur, gap = R.barcode_align_R(port, enc, seqs.seq_strings(), seqs.qual_strings(), barcodes)
    assert out["barcode"].tolist() == cid and np.array_equal(out["score"], np.array(cur)) and np.array_equal(out["gap"], np.array(gap))
    assert np.mean(out["barcode"] - 1 == pick) > 0.95
    one = api.barcodeAlign(seqs, barcodes[:1])
    assert np.all(np.isinf(one["gap"])) and set(one["barcode"].tolist()) == {1}
    # barcodes of different lengths take the reference's own loop
    mixed = barcodes[:3] + [barcodes[3][:20]]
    out = api.barcodeAlign(seqs, mixed)
    cid, cur, gap = R.barcode_align_R(port, enc, seqs.seq_strings(), seqs.qual_strings(), mixed)
    assert out["barcode"].tolist() == cid and np.array_equal(out["score"], np.array(cur))


def test_extract_subseq(port, enc):
    """R/extractSubseq.R: arbitrary sub-ranges by re-alignment; the stored scores must be reproduced exactly."""
    from sarlacc_b200 import api, synth
    from oracle import r_level as R
    reads = synth.mock_reads(120, VIGNETTE_A1, VIGNETTE_A2, seed=9)
    aligned = api.adaptorAlign(VIGNETTE_A1, VIGNETTE_A2, reads, number=50)
    sub1 = {"starts": np.array([1, 17, 30]), "ends": np.array([16, 28, 70])}
    sub2 = {"starts": np.array([5]), "ends": np.array([10])}
    out = api.extractSubseq(aligned, sub1, sub2, number=37)
    seqs, quals = reads.seq_strings(), reads.qual_strings()
    front, back = R.get_front_and_back(seqs, quals, 250)
    rev = aligned["reversed"]
    w1 = [back[i] if rev[i] else front[i] for i in range(len(seqs))]
    w2 = [front[i] if rev[i] else back[i] for i in range(len(seqs))]
    e1 = R.align_and_extract(port, enc, VIGNETTE_A1, w1, 5, 1, sub1["starts"].tolist(), sub1["ends"].tolist())
    e2 = R.align_and_extract(port, enc, VIGNETTE_A2, w2, 5, 1, sub2["starts"].tolist(), sub2["ends"].tolist())
    for k in range(3):
        assert out["adaptor1"]["Sub%d" % (k + 1)].seq_strings() == [e["subseq"][k][0] for e in e1]
    assert out["adaptor2"]["Sub1"].seq_strings() == [e["subseq"][0][0] for e in e2]
    # the barcode section asked for again equals what adaptorAlign stored
    assert out["adaptor1"]["Sub2"].seq_strings() == aligned["adaptor1"]["subseq"]["Sub1"].seq_strings()
    only2 = api.extractSubseq(aligned, subseq2=sub2)
    assert "adaptor1" not in only2 and only2["adaptor2"]["Sub1"].seq_strings() == out["adaptor2"]["Sub1"].seq_strings()
    # a tampered score is caught like in the reference
    aligned["adaptor1"]["score"] = aligned["adaptor1"]["score"] + 1.0
    with pytest.raises(RuntimeError, match="score mismatch from 'aligned' for adaptor 1"):
        api.extractSubseq(aligned, sub1)


def test_tune_alignment(port, enc):
    """tests/testthat/test-tuning.R:26-42 (the pairwiseAlignment comparison is replaced by the oracle)."""
    from sarlacc_b200 import api, ReadSet
    from oracle import r_level as R
    a1, a2 = "CGTACGACGAT", "TCGAGCGTTAC"
    reads = ["CGTACGACGATGACTGATCGATCGTAGTTCATCGACGATGTAACGCTCGA", "CGTCGACGATGACTGATCGATCGTAGTTCATCGACGATGTAACGCTCGA",
             "CGTACGACGATGACTGATCGATCGTAGTTCATCGACGATGTAACGCCGA", "CGTACGACGATGACTGATCGATCGTAGTTCATCGACGATGTAACGCTCGA",
             "CGCTACGACGATGACTGATCGATCGTAGTTCATCGACGATGTAACGGTCGA", "GTACGACGATTTACTGATCGATCGTAGTTCATCGACGATGTAACGCTCGA",
             "CGTCGACGATGACTGGGCGATCGTAGTTCATCGACGATGTAACGCTCGA", "TACGACGATGACTGATCATCGTAAAAATTCATCGATGTAACGCCGA",
             "CGTACGACGATGACTGATCGATCGTAGTTCACCCCGATGTAACGCTCGA", "CGCTACGACGATGACTGCGATCGTAGTTCATAAAAATGTAACGGTCGA"]
    rs = ReadSet.from_strings(reads, ["~" * len(r) for r in reads], ["READ_%d" % (i + 1) for i in range(len(reads))])
    out = api.tuneAlignment(a1, a2, rs, gapOp_range=(4, 5), gapExt_range=(1, 2))
    go, ge = out["parameters"]["gapOpening"], out["parameters"]["gapExtension"]
    assert go in (4, 5) and ge in (1, 2)
    assert out["scores"]["reads"].min() > out["scores"]["scrambled"].max()
    quals = ["~" * len(r) for r in reads]
    front, back = R.get_front_and_back(reads, quals, 200)
    sc = lambda w, a: port.align_score_only([x[0] for x in w], [x[1] for x in w], enc, go, ge, a)
    _, final = R.resolve_strand(sc(front, a1), sc(back, a2), sc(back, a1), sc(front, a2))
    assert np.array_equal(out["scores"]["reads"], np.array(final))
    none = api.tuneAlignment(a1, a2, ReadSet.empty())
    assert none["parameters"] == {"gapOpening": None, "gapExtension": None} and len(none["scores"]["reads"]) == 0


def test_quality_align(port, enc):
    from sarlacc_b200 import api, ReadSet
    from conftest import random_windows
    rng = np.random.default_rng(8)
    ref = "aaggaattaaggccttacgt"
    seqs, quals = random_windows(rng, 120, ref.upper(), 5, 40)
    out = api.qualityAlign(ReadSet.from_strings(seqs, quals), ref, gapOpening=4, gapExtension=1)
    exp = port.general_align(seqs, quals, enc, 4, 1, ref.upper())
    assert np.array_equal(out["score"], exp[0]) and np.array_equal(out["edit"], exp[1])
    assert out["reference"].tolist() == exp[2] and out["query"].tolist() == exp[3]
    # batching invariance (tests/testthat/test-general-align.R:81-93)
    half = api.qualityAlign(ReadSet.from_strings(seqs[:60], quals[:60]), ref, gapOpening=4, gapExtension=1)
    assert np.array_equal(half["score"], out["score"][:60])


def test_full_size_properties(port, enc):
    """At production scale the oracle cannot check everything; check size-independent properties on 200k reads
    (2 x 200k windows x 4 alignments) and a strided sample against the oracle."""
    from sarlacc_b200 import native, synth
    n = 200000
    front, back, widths, flips = synth.mock_windows_device(n, VIGNETTE_A1, VIGNETTE_A2, seed=2000)
    ss, se = [16, 42], [28, 46]
    a = native.adaptor_align(front, enc, 5, 1, VIGNETTE_A1, ss, se)
    # (1) score-only pass, resident pass and host-buffer pass agree bit for bit (batching / path invariance)
    assert np.array_equal(native.adaptor_align_score_only(front, enc, 5, 1, VIGNETTE_A1), a[0])
    r = native.Resident(front, enc)
    r.align(r.MODE_TRACE_LOCAL, 5, 1, VIGNETTE_A1, ss, se)
    b = r.fetch()
    r.close()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    for s in range(2):
        assert np.array_equal(a[3][s], b[3][s]) and np.array_equal(a[4][s], b[4][s])
    # (2) coordinate invariants: 1 <= start <= end <= 250 when aligned; sections lie inside the window and are ordered
    ok = a[1] > 0
    assert np.all(a[1][ok] <= a[2][ok]) and np.all(a[2] <= 250)
    for s in range(2):
        assert np.all(a[3][s] >= 1) and np.all(a[3][s] - 1 + a[4][s] <= 250) and np.all(a[4][s] >= 0)
    assert np.all(a[3][0] - 1 + a[4][0] <= a[3][1] - 1 + np.maximum(a[4][1], 0) + 250 * 0 + 250)  # barcode run ends before the window does
    # (3) un-flipped reads carry adaptor1 at the front: high score, start near 1, 12-base barcode, 4-base UMI mostly
    fwd = ~flips
    assert np.median(a[0][fwd]) > 40 and np.median(a[1][fwd]) == 1
    assert np.mean(a[4][0][fwd] == 12) > 0.7 and np.mean(a[4][1][fwd] == 4) > 0.8
    assert np.median(a[0][flips]) < 10
    # (4) strided samples of all four alignments of .align_AA_internal against the oracle
    got = {"a1_front": a, "a2_back": native.adaptor_align(back, enc, 5, 1, VIGNETTE_A2),
           "a1_back": native.adaptor_align(back, enc, 5, 1, VIGNETTE_A1, ss, se), "a2_front": native.adaptor_align(front, enc, 5, 1, VIGNETTE_A2)}
    for k, (key, rs, ad, sec) in enumerate((("a1_front", front, VIGNETTE_A1, True), ("a2_back", back, VIGNETTE_A2, False),
                                            ("a1_back", back, VIGNETTE_A1, True), ("a2_front", front, VIGNETTE_A2, False))):
        idx = np.arange(k * 13, n, 97)
        sub = rs[idx]
        exp = port.adaptor_align(sub.seq_strings(), sub.qual_strings(), enc, 5, 1, ad, ss if sec else [], se if sec else [], nthreads=8)
        g = got[key]
        assert np.array_equal(g[0][idx], exp[0]) and np.array_equal(g[1][idx], exp[1]) and np.array_equal(g[2][idx], exp[2]), key
        for s in range(2 if sec else 0):
            assert np.array_equal(g[3][s][idx], exp[3][s]) and np.array_equal(g[4][s][idx], exp[4][s]), key


def test_multi_device_sharding_in_one_process(port, enc):
    """sarlacc_set_devices: contiguous read-index shards over every visible GPU, results identical to one device
    (and to the oracle).  With a single GPU this degenerates to listing device 0 twice-over shards."""
    import ctypes
    from sarlacc_b200 import native, _lib
    from conftest import random_windows
    ndev = _lib.lib.sarlacc_device_count()
    rng = np.random.default_rng(21)
    seqs, quals = random_windows(rng, 1500, VIGNETTE_A2, 5, 120)
    exp = port.adaptor_align(seqs, quals, enc, 5, 1, "AAGGCCTTNNNNCGACTCATGAA", [8], [12], nthreads=4)
    devs = list(range(ndev))
    try:
        _lib.check(_lib.lib.sarlacc_set_devices((ctypes.c_int * len(devs))(*devs), len(devs)))
        got = native.adaptor_align((seqs, quals), enc, 5, 1, "AAGGCCTTNNNNCGACTCATGAA", [8], [12])
        assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1]) and np.array_equal(got[2], exp[2])
        assert np.array_equal(got[3][0], exp[3][0]) and np.array_equal(got[4][0], exp[4][0])
        bid, best, nxt = native.barcode_align_multi((seqs, quals), enc, 5, 1, ["AAGGCCTTTTCCGACTCATGAA", "AAGGCCTTTTCCGACTCATGTT"])
        e0 = port.align_score_only(seqs, quals, enc, 5, 1, "AAGGCCTTTTCCGACTCATGAA", local=False)
        e1 = port.align_score_only(seqs, quals, enc, 5, 1, "AAGGCCTTTTCCGACTCATGTT", local=False)
        assert np.array_equal(best, np.maximum(e0, e1)) and np.array_equal(bid, np.where(e1 > e0, 2, 1))
        with pytest.raises(native.SarlaccError, match="device index out of range"):
            _lib.check(_lib.lib.sarlacc_set_devices((ctypes.c_int * 1)(ndev + 3), 1))
    finally:
        _lib.check(_lib.lib.sarlacc_set_devices((ctypes.c_int * 1)(0), 1))


def test_long_reads_and_long_references(port, enc):
    """Rows stream through the wavefront, so read length is unbounded; references up to 256 columns use it too,
    longer ones the literal kernel.  (qualityAlign-style global alignment of kilobase sequences.)"""
    from sarlacc_b200 import native
    rng = np.random.default_rng(33)
    ref = "".join(rng.choice(list("ACGT"), size=200))
    seqs, quals = [], []
    for _ in range(40):
        n = int(rng.integers(1500, 3000))
        s = rng.choice(list("ACGT"), size=n).tolist()
        p = int(rng.integers(0, n - 250))
        s[p:p + 200] = list(ref)
        seqs.append("".join(s))
        quals.append("".join(chr(33 + int(q)) for q in rng.integers(5, 40, size=n)))
    got = native.adaptor_align((seqs, quals), enc, 5, 1, ref, [10, 100], [20, 150])
    exp = port.adaptor_align(seqs, quals, enc, 5, 1, ref, [10, 100], [20, 150], nthreads=8)
    for a, b in zip(got[:3], exp[:3]):
        assert np.array_equal(a, b)
    for s in range(2):
        assert np.array_equal(got[3][s], exp[3][s]) and np.array_equal(got[4][s], exp[4][s])
    g = native.general_align((seqs[:10], quals[:10]), enc, 4, 1, ref)
    e = port.general_align(seqs[:10], quals[:10], enc, 4, 1, ref)
    assert np.array_equal(g[0], e[0]) and np.array_equal(g[1], e[1]) and g[2] == e[2] and g[3] == e[3]


def test_tune_alignment_batched_equals_the_reference_loop(enc):
    """The batched grid (one chunk, scores kept on the device, .tied_overlap there) picks the same point and returns the
    same score vectors as the reference's loop over the four score-only calls + .resolve_strand + .tied_overlap on the
    host (R/tuneAlignment.R:54-72) -- on a random sample of a larger read set."""
    from sarlacc_b200 import api, native, synth
    reads = synth.mock_reads(400, VIGNETTE_A1, VIGNETTE_A2, seed=5)
    grid_args = dict(gapOp_range=(4, 6), gapExt_range=(1, 2), tolerance=150, number=150, seed=4)
    out = api.tuneAlignment(VIGNETTE_A1, VIGNETTE_A2, reads, **grid_args)
    sample = api._sample_reads(reads, 150, 150, 4)
    assert len(sample) == 150 and len(set(sample.names)) == 150
    grid = [(go, ge) for go in (4, 5, 6) for ge in (1, 2)]
    ref = api._tune_alignment_host(sample, VIGNETTE_A1, VIGNETTE_A2, 150, grid, api._create_encoding_vector("PhredQuality"), 4)
    assert out["parameters"] == ref["parameters"]
    assert np.array_equal(out["scores"]["reads"], ref["scores"]["reads"]) and np.array_equal(out["scores"]["scrambled"], ref["scores"]["scrambled"])
    assert out["scores"]["reads"].mean() > out["scores"]["scrambled"].mean() + 20
    # .tied_overlap on the device == the R expression, ties included (tests/testthat/test-tuning.R:53-59)
    rng = np.random.default_rng(1)
    real, fake = np.round(rng.normal(5, 2, 5000), 1), np.round(rng.normal(4, 2, 3000), 1)
    assert native.tied_overlap(real, fake) == api._tied_overlap(real, fake)
    assert native.tied_overlap(np.array([1.0, 2.0]), np.array([1.0, 2.0])) == api._tied_overlap([1.0, 2.0], [1.0, 2.0]) == 0.5
    first = api.tuneAlignment(VIGNETTE_A1, VIGNETTE_A2, reads, sample="first", **grid_args)
    assert first["parameters"]["gapOpening"] in (4, 5, 6) and len(first["scores"]["reads"]) == 150


def test_bases_sent_as_four_bit_codes(enc, monkeypatch):
    """With SARLACC_PACK_SEQ=1 a both-ends job sends the bases as 4-bit codes: same results as the plain bytes, for ragged windows of odd lengths with N, IUPAC and lower-case bases (all "not ACGT" to the packer),
    pageable and pinned pools, and fewer bytes on the link."""
    import torch
    from conftest import random_windows, stable_seed
    from sarlacc_b200 import native, ReadSet
    rng = np.random.default_rng(stable_seed("nibbles"))
    n = 4000
    fs, fq = random_windows(rng, n, VIGNETTE_A1, 1, 251, alphabet="ACGTACGTACGTNRacgt")
    bs, bq = random_windows(rng, n, VIGNETTE_A2, 0, 130, alphabet="ACGTACGTACGTNYt")
    bs = [b if len(b) else "A" for b in bs]
    bq = [q if len(q) else "I" for q in bq]
    front, back = ReadSet.from_strings(fs, fq), ReadSet.from_strings(bs, bq)
    widths = (front.width() + back.width() + 100).astype(np.int32)
    s1, e1 = [16, 42], [28, 46]
    monkeypatch.setenv("SARLACC_CHUNK", "1100")
    results = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("SARLACC_PACK_SEQ", mode)
        results[mode] = native.adaptor_align_windows(front, back, enc, 5, 1, VIGNETTE_A1, VIGNETTE_A2, (s1, e1), ((), ()), read_width=widths)
        results[mode + "bytes"] = native.last_pair_timing()["upload_bytes"]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()   # noqa: E731
    pf = ReadSet(pin(front.seq_pool), front.seq_off, pin(front.qual_pool), front.qual_off, front.names)
    pb = ReadSet(pin(back.seq_pool), back.seq_off, pin(back.qual_pool), back.qual_off, back.names)
    results["pinned"] = native.adaptor_align_windows(pf, pb, enc, 5, 1, VIGNETTE_A1, VIGNETTE_A2, (s1, e1), ((), ()), read_width=widths)
    a = results["0"]
    for other in (results["1"], results["pinned"]):
        assert np.array_equal(a[0], other[0])
        for x, y in ((a[1], other[1]), (a[2], other[2])):
            for k in range(3):
                assert np.array_equal(x[k], y[k])
            for k in range(len(x[3])):
                assert np.array_equal(x[3][k], y[3][k]) and np.array_equal(x[4][k], y[4][k])
    nbases = int(front.seq_off[-1] + back.seq_off[-1])
    assert results["0bytes"] - results["1bytes"] >= nbases // 2 - 8       # half the sequence bytes stayed at home
