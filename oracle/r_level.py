"""TEST INFRASTRUCTURE ONLY -- loop-based restatement of the R drivers around the hot path, on top of the CPU
oracle (oracle.py).  Plain Python strings and loops, written line by line after the R sources so that it is
independent of the vectorised product code in sarlacc_b200/api.py; meant for small cases.

    setup_subseqs        R/adaptorAlign.R:136-143
    get_front_and_back   R/adaptorAlign.R:86-95
    resolve_strand       R/adaptorAlign.R:112-122
    align_and_extract    R/adaptorAlign.R:150-176
    adaptor_align_R      R/adaptorAlign.R:7-78,178-199
    compute_threshold    R/getAdaptorThresholds.R:94-103
    adaptor_thresholds_R R/getAdaptorThresholds.R:6-66,105-128 (scrambled windows supplied by the caller)
    barcode_align_R      R/barcodeAlign.R:4-40
    tied_overlap         R/tuneAlignment.R:78-86
"""
import math
import re

_COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "M": "K", "K": "M", "R": "Y", "Y": "R", "V": "B", "B": "V",
         "H": "D", "D": "H", "W": "W", "S": "S", "N": "N"}


def revcomp(s):
    return "".join(_COMP.get(c, c) for c in reversed(s))


def setup_subseqs(adaptor):
    starts, ends = [], []
    for m in re.finditer("[^ACTG]+", adaptor):
        starts.append(m.start() + 1)
        ends.append(m.start() + 1 + (m.end() - m.start()) - 1)
    return starts, ends


def get_front_and_back(seqs, quals, tolerance):
    front, back = [], []
    for s, q in zip(seqs, quals):
        tol = min(tolerance, len(s))
        front.append((s[:tol], q[:tol]))
        es, eq = s[len(s) - tol:], q[len(q) - tol:]
        back.append((revcomp(es), eq[::-1]))
    return front, back


def resolve_strand(start, end, rc_start, rc_end):
    rev, final = [], []
    for a, b, c, d in zip(start, end, rc_start, rc_end):
        f = max(a, 0) + max(b, 0)
        r = max(c, 0) + max(d, 0)
        rev.append(f < r)
        final.append(r if f < r else f)
    return rev, final


def align_and_extract(O, enc, adaptor, windows, go, ge, starts, ends):
    seqs = [w[0] for w in windows]
    quals = [w[1] for w in windows]
    score, start, end, sst, swd = O.adaptor_align(seqs, quals, enc, go, ge, adaptor, [x - 1 for x in starts], ends)
    rows = []
    for i in range(len(windows)):
        subs = []
        for k in range(len(starts)):
            a, w = int(sst[k][i]), int(swd[k][i])
            subs.append((seqs[i][a - 1:a - 1 + w], quals[i][a - 1:a - 1 + w]))
        rows.append({"score": float(score[i]), "start": int(start[i]), "end": int(end[i]), "subseq": subs})
    return rows


def adaptor_align_R(O, enc, adaptor1, adaptor2, seqs, quals, tolerance=250, go=5, ge=1):
    adaptor1, adaptor2 = adaptor1.upper(), adaptor2.upper()
    s1, e1 = setup_subseqs(adaptor1)
    s2, e2 = setup_subseqs(adaptor2)
    front, back = get_front_and_back(seqs, quals, tolerance)
    cs = align_and_extract(O, enc, adaptor1, front, go, ge, s1, e1)
    ce = align_and_extract(O, enc, adaptor2, back, go, ge, s2, e2)
    rs = align_and_extract(O, enc, adaptor1, back, go, ge, s1, e1)
    re_ = align_and_extract(O, enc, adaptor2, front, go, ge, s2, e2)
    rev, _ = resolve_strand([x["score"] for x in cs], [x["score"] for x in ce],
                            [x["score"] for x in rs], [x["score"] for x in re_])
    out = []
    for i in range(len(seqs)):
        a1 = rs[i] if rev[i] else cs[i]
        a2 = dict(re_[i] if rev[i] else ce[i])
        width = len(seqs[i])
        a2["start"] = width - a2["start"] + 1
        a2["end"] = width - a2["end"] + 1
        out.append({"read.width": width, "adaptor1": a1, "adaptor2": a2, "reversed": rev[i]})
    return out


def find_interval(x, vec):
    """R's findInterval(x, vec) for sorted vec: number of elements <= x."""
    n = 0
    for v in vec:
        if v <= x:
            n += 1
        else:
            break
    return n


def compute_threshold(real, scrambled, error):
    real = sorted(real)
    scrambled = sorted(scrambled)
    for k, r in enumerate(real, start=1):
        num = len(scrambled) - find_interval(r, scrambled)
        den = len(real) - k
        if den == 0:
            fdr = float("nan") if num == 0 else float("inf")
        else:
            fdr = num / den
        if not math.isnan(fdr) and fdr <= error:
            return r
    return float("nan")


def adaptor_thresholds_R(O, enc, adaptor1, adaptor2, scr_front, scr_back, real1, real2, go, ge, error=0.01):
    """scr_front/scr_back: lists of (seq, qual) scrambled windows (the R RNG is not reproducible; identity is
    defined from the scrambled windows onwards, SURVEY 7 'hard parts')."""
    def sc(w, a):
        return O.align_score_only([x[0] for x in w], [x[1] for x in w], enc, go, ge, a)
    S, E, RS, RE = sc(scr_front, adaptor1), sc(scr_back, adaptor2), sc(scr_back, adaptor1), sc(scr_front, adaptor2)
    rev, _ = resolve_strand(S, E, RS, RE)
    s1 = [RS[i] if rev[i] else S[i] for i in range(len(rev))]
    s2 = [RE[i] if rev[i] else E[i] for i in range(len(rev))]
    return compute_threshold(real1, s1, error), compute_threshold(real2, s2, error), s1, s2


def barcode_align_R(O, enc, seqs, quals, barcodes, go=5, ge=1):
    n = len(seqs)
    cur = [-math.inf] * n
    nxt = [-math.inf] * n
    cid = [0] * n
    for b, bc in enumerate(barcodes, start=1):
        scores = O.align_score_only(seqs, quals, enc, go, ge, bc, local=False)
        for i in range(n):
            s = scores[i]
            if s > cur[i]:
                cid[i] = b
                nxt[i] = cur[i]
                cur[i] = s
            elif s > nxt[i]:
                nxt[i] = s
    return cid, cur, [c - x for c, x in zip(cur, nxt)]


def tied_overlap(real, fake):
    fake = sorted(fake)
    tot = 0.0
    for r in real:
        upper = sum(1 for f in fake if f <= r)
        lower = sum(1 for f in fake if f < r)
        tot += (upper + lower) / 2.0
    return tot / (len(real) * len(fake))
