"""Offline tuning helper: compile kernels.cu for one geometry with extra -D flags and report, for the row loop of
wf_forward<C,trace>, the static instruction count per DP cell (largest loop body containing SHFL and DSETP, divided by
rows-per-trip x C), registers and spills.  No GPU needed.
usage: python tools/variants.py C "flags" ["flags" ...]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C = int(sys.argv[1])
SRC = os.path.join(ROOT, "sarlacc_b200", "csrc")
os.makedirs(os.path.join(ROOT, "scratch"), exist_ok=True)


def analyse(cubin, trace):
    out = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    lines = out.splitlines()
    pat = "wf_forwardILi%dELb%dEEE" % (C, trace)
    st = next(i for i, l in enumerate(lines) if "Function :" in l and pat in l)
    en = next((i for i in range(st + 1, len(lines)) if "Function :" in lines[i]), len(lines))
    ins = []
    for l in lines[st:en]:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    best = None
    for addr, txt in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", txt)
        if m and int(m.group(1), 16) < addr:
            body = [t for a, t in ins if int(m.group(1), 16) <= a <= addr]
            n_shfl = sum("SHFL" in t for t in body)
            if n_shfl >= 4 and any("DSETP" in t for t in body):
                rows = n_shfl // 4
                per = len(body) / (rows * C)
                if best is None or per < best[0]:
                    best = (per, rows, body)
    cnt = collections.Counter()
    for t in best[2]:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        cnt[t.split()[0].split(".")[0]] += 1
    rows = best[1]
    top = ", ".join("%s %.1f" % (k, v / (rows * C)) for k, v in cnt.most_common(9))
    return best[0], rows, top


for flags in sys.argv[2:]:
    tag = re.sub(r"[^A-Za-z0-9]+", "_", flags)[:60] or "base"
    cubin = os.path.join(ROOT, "scratch", "var_%s.cubin" % tag)
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "--fmad=false", "-std=c++17",
           "-I" + os.path.join(ROOT, "include"), "-I" + SRC, "-DSARLACC_ONLY_C=%d" % C, "-Xptxas", "-v", "-cubin", "-o", cubin,
           os.path.join(SRC, "kernels.cu")] + flags.split()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(flags, "FAILED", r.stderr[-400:])
        continue
    regs = {}
    for block in r.stderr.split("Compiling entry function")[1:]:
        mm = re.search(r"wf_forwardILi(\d+)ELb([01])E", block)
        sp = re.search(r"(\d+) bytes spill stores", block)
        rg = re.search(r"Used (\d+) registers", block)
        if mm and sp and rg:
            if int(mm.group(1)) == C:
                regs[int(mm.group(2))] = (rg.group(1), sp.group(1))
    for trace in (1, 0):
        per, rows, top = analyse(cubin, trace)
        print("%-46s trace=%d  %.2f instr/cell (x%d rows)  regs=%s spill=%s | %s" % (flags or "(base)", trace, per, rows, regs[trace][0], regs[trace][1], top))
