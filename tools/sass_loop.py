"""Static instruction mix of the unmasked DP row loop of one forward-kernel variant (cuobjdump -sass, no GPU).
usage: python tools/sass_loop.py C trace [kernel=wf_forward2] [solo=0|1] [--dump] [--lib path]
The row loop is the smallest backward-branch body that holds at least 4*C DSETP (one row of C cells has 4 FP64 compares)."""
import collections
import re
import subprocess
import sys

_CACHE = {}


def _sass(lib):
    if lib not in _CACHE:
        _CACHE[lib] = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
    return _CACHE[lib]


def analyse(lib, name, C, trace, solo=0):
    """(instructions per cell, rows per trip, Counter of mnemonics, loop body) of kernel name<C, trace[, solo]>."""
    C = int(C)
    lines = _sass(lib)
    pat = "%sILi%dELb%dE" % (name, C, int(trace)) + ("Lb%dE" % int(solo) if name == "wf_forward2" else "") + "EE"
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
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr:
                body = [t for a, t in ins if tgt <= a <= addr]
                if sum("DSETP" in t for t in body) >= 4 * C:
                    if best is None or len(body) < len(best):
                        best = body
    cnt = collections.Counter()
    for t in best:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        cnt[t.split()[0].split(".")[0]] += 1
    rows = max(1, round(cnt["DSETP"] / (4.0 * C)))
    return len(best) / (rows * C), rows, cnt, best


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    C, tr = args[0], args[1]
    name = args[2] if len(args) > 2 else "wf_forward2"
    solo = args[3] if len(args) > 3 else "0"
    lib = "sarlacc_b200/libsarlacc_b200.so"
    if "--lib" in sys.argv:
        lib = sys.argv[sys.argv.index("--lib") + 1]
        args = [a for a in args if a != lib]
    per, rows, cnt, body = analyse(lib, name, C, tr, solo)
    print("%s<C=%s, trace=%s, solo=%s>: loop instructions %d = %d rows x %s columns -> %.2f per cell" %
          (name, C, tr, solo, len(body), rows, C, per))
    print(sorted(cnt.items(), key=lambda kv: -kv[1]))
    if "--dump" in sys.argv:
        print("\n".join(body))
