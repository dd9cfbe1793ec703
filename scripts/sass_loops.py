"""Static per-loop summary of a kernel's SASS: instructions, FP64 instructions and the sum of the stall counts that ptxas
encoded in the control bits (the cycles one warp needs for an iteration if nothing else runs and every scoreboard is ready).
Loops are found from backward branches.  Usage: python scripts/sass_loops.py <obj> <mangled-name-pattern> [min_instr]"""
import re
import subprocess
import sys
from collections import Counter

obj, pat = sys.argv[1], sys.argv[2]
min_instr = int(sys.argv[3]) if len(sys.argv) > 3 else 60
names = subprocess.run("cuobjdump -elf %s | grep -o '_ZN2mb[A-Za-z0-9_]*' | grep -E '%s' | grep -v _param_ | sort -u | head -1" % (obj, pat),
                       shell=True, capture_output=True, text=True).stdout.split()
txt = subprocess.run(["cuobjdump", "-sass", "-fun", names[0], obj], capture_output=True, text=True).stdout.splitlines()
ins = []  # (addr, stall, text)
i = 0
while i < len(txt):
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", txt[i])
    if m and i + 1 < len(txt):
        m2 = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", txt[i + 1])
        if m2:
            w = int(m2.group(1), 16)
            ins.append((int(m.group(1), 16), (w >> (105 - 64)) & 0xf, m.group(2).strip()))
        i += 2
    else:
        i += 1
print("%s: %d instructions" % (names[0][:90], len(ins)))
loops = []
for a, st, s in ins:
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", s)
    if m and int(m.group(1), 16) < a:
        loops.append((int(m.group(1), 16), a))
for t, a in loops:
    body = [(ad, st, s) for ad, st, s in ins if t <= ad <= a]
    if len(body) < min_instr:
        continue
    c = Counter()
    for _, _, s in body:
        p = s.split()
        op = p[1] if p[0].startswith("@") else p[0]
        c[op.split(".")[0]] += 1
    fp = c["DFMA"] + c["DMUL"] + c["DADD"]
    print("loop %05x-%05x: %4d instr, fp64 %3d, stall sum %4d | BRA %d IMAD %d LDCU %d U* %d LDS %d LDL %d STL %d" % (
        t, a, len(body), fp, sum(st for _, st, _ in body), c["BRA"], c["IMAD"], c["LDCU"],
        sum(v for k, v in c.items() if k.startswith("U")), c["LDS"], c["LDL"], c["STL"]))
