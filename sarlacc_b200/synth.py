"""mockReads-style synthetic reads (R/mockReads.R), re-implemented with numpy's counter-based Philox so
that every read depends only on (seed, read index) -- the same data whatever the chunking or GPU count.

Molecule = adaptor1 (first N-run <- barcode, second N-run <- random UMI) + uniform ACGT insert +
revcomp(adaptor2) (R/mockReads.R:58-64); per read each base is substituted w.p. sub_rate by a uniform
base (:73-74), then w.p. indel_rate replaced by 0 or 2..max_insert copies of itself (:77-79); qualities
are iid Phred of U(0, sub_rate+indel_rate) (:82); half of the reads are reverse-complemented (:91-92).

Only `tolerance` bases from either end are ever aligned (R/adaptorAlign.R:86-95), so `mock_windows`
materialises just those two windows per read (plus the read width); `mock_reads` materialises whole
reads and is meant for small cases (configs[0]).
"""
import numpy as np

from .reads import ReadSet, _COMP

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _n_runs(adaptor):
    import re
    return [(m.start(), m.end()) for m in re.finditer("N+", adaptor)]


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


def _mutate_take(rng, mol, take, sub_rate, indel_rate, max_insert):
    """Mutate each row of `mol` (n, P) like mockReads and return the first `take` bases of every mutated row
    as an (n, take) matrix (rows are long enough by construction)."""
    n, P = mol.shape
    sub = rng.random((n, P)) < sub_rate
    mol = mol.copy()
    mol[sub] = _ACGT[rng.integers(0, 4, size=int(sub.sum()))]
    counts = np.ones((n, P), dtype=np.int64)
    ind = rng.random((n, P)) < indel_rate
    choices = np.array([0] + list(range(2, max_insert + 1)), dtype=np.int64)
    counts[ind] = choices[rng.integers(0, len(choices), size=int(ind.sum()))]
    tot = counts.sum(axis=1)
    if np.any(tot < take):
        raise ValueError("molecule prefix too short for the requested window")
    flat = np.repeat(mol.reshape(-1), counts.reshape(-1))
    off = np.zeros(n, dtype=np.int64)
    np.cumsum(tot[:-1], out=off[1:])
    idx = off[:, None] + np.arange(take, dtype=np.int64)[None, :]
    return flat[idx]


def _phred_quals(rng, shape, max_err):
    p = rng.random(shape) * max_err
    p = np.maximum(p, 1e-12)
    q = np.clip(np.rint(-10.0 * np.log10(p)), 0, 93).astype(np.uint8)
    return (q + 33).astype(np.uint8)


def mock_windows(n, adaptor1, adaptor2, tolerance=250, seed=2000, insert_len=4908, barcodes=None,
                 sub_rate=0.05, indel_rate=0.01, max_insert=5, first_index=0, block=50000):
    """Front and back windows (as .get_front_and_back would cut them, the back one reverse-complemented) of
    n synthetic reads.  Returns (front ReadSet, back ReadSet, read widths int64[n], flipped bool[n]).
    Read i of a run is generated from Philox(key=seed, counter block = first_index+i)-derived streams, so
    shards generated with different `first_index` tile one big data set."""
    a1 = adaptor1.upper()
    a2 = adaptor2.upper()
    W = int(tolerance)
    P = W + 60   # molecule prefix long enough to survive deletions
    fronts, backs, widths, flips = [], [], [], []
    for b0 in range(0, n, block):
        m = min(block, n - b0)
        rng = np.random.Generator(np.random.Philox(key=seed, counter=[0, 0, 0, first_index + b0]))
        head = np.concatenate([_fill_adaptor(rng, a1, m, barcodes), _ACGT[rng.integers(0, 4, size=(m, max(0, P - len(a1))))]], axis=1)[:, :P]
        tail = np.concatenate([_fill_adaptor(rng, a2, m, None), _ACGT[rng.integers(0, 4, size=(m, max(0, P - len(a2))))]], axis=1)[:, :P]
        hw = _mutate_take(rng, head, W, sub_rate, indel_rate, max_insert)
        tw = _mutate_take(rng, tail, W, sub_rate, indel_rate, max_insert)
        hq = _phred_quals(rng, (m, W), sub_rate + indel_rate)
        tq = _phred_quals(rng, (m, W), sub_rate + indel_rate)
        flip = rng.random(m) < 0.5
        # read width: mutated length of adaptor1 + insert + adaptor2
        M = len(a1) + insert_len + len(a2)
        cnt = rng.binomial(M, indel_rate, size=m)
        k = rng.multinomial(cnt, [1.0 / max_insert] * max_insert)
        delta = k @ np.array([-1] + list(range(1, max_insert)), dtype=np.int64)
        widths.append(M + delta)
        f_seq = np.where(flip[:, None], tw, hw)
        b_seq = np.where(flip[:, None], hw, tw)
        f_q = np.where(flip[:, None], tq, hq)
        b_q = np.where(flip[:, None], hq, tq)
        fronts.append((f_seq, f_q))
        backs.append((b_seq, b_q))
        flips.append(flip)

    def build(parts):
        seq = np.concatenate([p[0] for p in parts]).reshape(-1)
        qual = np.concatenate([p[1] for p in parts]).reshape(-1)
        off = np.arange(n + 1, dtype=np.int64) * W
        return ReadSet(seq, off, qual, off, None)

    return build(fronts), build(backs), np.concatenate(widths), np.concatenate(flips)


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
