"""Print the instruction mix of the innermost loop (the DP row loop) of one wf_forward variant.
usage: python tools/sass_loop.py C trace [--dump]"""
import re, subprocess, sys, collections
C, tr = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", "sarlacc_b200/libsarlacc_b200.so"], capture_output=True, text=True).stdout
name = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else "wf_forward2"
pat = "%sILi%sELb%sEEE" % (name, C, tr)
lines = out.splitlines()
st = next(i for i, l in enumerate(lines) if "Function :" in l and pat in l)
en = next((i for i in range(st + 1, len(lines)) if "Function :" in lines[i]), len(lines))
ins = []
for l in lines[st:en]:
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
# find backward branches; the innermost loop = the backward branch with the largest body containing SHFL and DSETP
best = None
for idx, (addr, txt) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", txt)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < addr:
            body = [t for a, t in ins if tgt <= a <= addr]
            if any("SHFL" in t for t in body) and any("DSETP" in t for t in body):
                if best is None or len(body) < len(best):
                    best = body
cnt = collections.Counter()
for t in best:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    cnt[t.split()[0].split(".")[0]] += 1
print("loop instructions: %d  (%.1f per cell at C=%s)" % (len(best), len(best) / int(C), C))
print(sorted(cnt.items(), key=lambda kv: -kv[1]))
if "--dump" in sys.argv:
    print("\n".join(best))
