"""Generates tests/golden/umi_vectors.json from the reference's OWN umi_group / fast_levdist_test / cluster_umis_test
(oracle/_ref/libsarlacc_umi_ref.so: six reference source files compiled verbatim, oracle/Makefile `umiref`).
Run in the container that has /root/reference:  python tests/golden/make_umi_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.umi import UmiRef  # noqa: E402
from umi_cases import CASES, make_case  # noqa: E402


def attempt(f, *a):
    try:
        return {"value": f(*a)}
    except RuntimeError as e:
        return {"error": str(e)}


def main():
    R = UmiRef()
    out = []
    for name, seed, kw, t1, t2 in CASES:
        u1, u2, groups = make_case(seed, **kw)
        out.append({
            "name": name, "seed": seed, "threshold1": t1, "threshold2": t2, "umi1": u1, "umi2": u2, "groups": groups,
            "levdist": attempt(R.levdist, u1, t1, True),
            "one": attempt(R.umi_group, u1, t1),
            "one_grouped": attempt(R.umi_group, u1, t1, None, None, groups),
            "two_grouped": attempt(R.umi_group, u1, t1, u2, t2, groups),
        })
    with open(os.path.join(HERE, "umi_vectors.json"), "w") as fh:
        json.dump({"generator": "tests/golden/make_umi_golden.py", "cases": out}, fh, separators=(",", ":"))
    print("wrote", len(out), "cases;", sum("error" in c[k] for c in out for k in ("one", "one_grouped", "two_grouped")), "error outcomes")


if __name__ == "__main__":
    main()
