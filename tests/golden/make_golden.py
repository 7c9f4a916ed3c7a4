"""Generates tests/golden/*.json from the reference's OWN code (oracle/_ref/libsarlacc_ref.so = the reference's
reference_align.cpp + quality_encoding.cpp compiled verbatim from /root/reference by oracle/Makefile).

Run here (the container that has /root/reference):  python tests/golden/make_golden.py
The fixtures travel with the repo; the GPU box only reads them.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.oracle import Oracle, phred_encoding  # noqa: E402
from conftest import VIGNETTE_A1, VIGNETTE_A2, random_windows  # noqa: E402


def hexf(a):
    return [float(x).hex() for x in a]


def main():
    O = Oracle("ref")
    enc = phred_encoding()
    cases = []
    rng = np.random.default_rng(20261017)
    specs = [
        ("vignette_a1_250", VIGNETTE_A1, 5, 1, 150, 250, 12, 40, 60),
        ("vignette_a2_250", VIGNETTE_A2, 5, 1, 150, 250, 12, 40, 60),
        ("vignette_a1_short", VIGNETTE_A1, 4, 1, 1, 90, 0, 60, 60),
        ("iupac_mix", "ACGTNNRYACGTVVACKM", 5, 1, 5, 70, 0, 40, 60),
        ("ties_uniform_q", "ACACACACNNNNNNACACAC", 1, 1, 30, 90, 20, 20, 60),
        ("frac_gaps", "AAGGCCTTTTCCGACTCATGAA", 2.5, 0.3, 10, 80, 0, 40, 40),
    ]
    for name, adaptor, go, ge, lo, hi, qlo, qhi, n in specs:
        alphabet = "AC" if name.startswith("ties") else "ACGT"
        seqs, quals = random_windows(rng, n, adaptor, lo, hi, qlo, qhi, alphabet=alphabet)
        seqs += [""]
        quals += [""]
        import re
        st = [m.start() for m in re.finditer("[^ACTG]+", adaptor)]
        en = [m.end() for m in re.finditer("[^ACTG]+", adaptor)]
        score, start, end, sst, swd = O.adaptor_align(seqs, quals, enc, go, ge, adaptor, st, en)
        gscore = O.align_score_only(seqs, quals, enc, go, ge, adaptor, local=False)
        cases.append({"name": name, "adaptor": adaptor, "go": go, "ge": ge, "seqs": seqs, "quals": quals,
                      "sec_starts": st, "sec_ends": en, "score": hexf(score), "start": start.tolist(), "end": end.tolist(),
                      "sec_start": sst.tolist(), "sec_width": swd.tolist(), "global_score": hexf(gscore)})
    # general_align strings
    ref = "AAGGAATTAAGGCCTTACGT"
    seqs, quals = random_windows(rng, 40, ref, 5, 40)
    score, edit, rs, qs = O.general_align(seqs, quals, enc, 4, 1, ref)
    general = {"reference": ref, "go": 4, "ge": 1, "seqs": seqs, "quals": quals, "score": hexf(score),
               "edit": edit.tolist(), "ref_aln": rs, "query_aln": qs}
    m, mm, off = O.cost_tables(enc)
    tables = {"match": [hexf(r) for r in m], "mismatch": [hexf(r) for r in mm], "offset": off.decode()}
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as fh:
        json.dump({"generator": "tests/golden/make_golden.py", "source": "oracle/_ref (reference C++ compiled verbatim)",
                   "encoding": "phred 0..93 offset 33", "adaptor_cases": cases, "general": general, "cost_tables": tables}, fh)
    print("wrote", os.path.join(HERE, "reference_vectors.json"))


if __name__ == "__main__":
    main()
