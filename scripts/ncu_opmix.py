"""Aggregate an `ncu --page source --csv --print-source sass` dump by opcode: executed warp-instructions and stall samples
per kernel.  Usage: python scripts/ncu_opmix.py gpurun_out/x_src.csv [top]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kern = None
hdr = None
acc = {}
for row in csv.reader(open(path)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        kern = row[1]
        acc[kern] = {"ops": defaultdict(lambda: [0, 0]), "n": 0, "samples": 0, "stalls": defaultdict(int)}
        hdr = None
        continue
    if row[0] == "Address":
        hdr = {h: i for i, h in enumerate(row)}
        continue
    if kern is None or hdr is None:
        continue
    src = row[hdr["Source"]].strip()
    toks = src.split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.rstrip(";")
    base = op.split(".")[0]
    if base in ("LDS", "STS", "LDG", "STG", "LDL", "STL", "LDC", "LDCU"):
        base = op  # keep the width
    ex = int(float(row[hdr["Instructions Executed"]] or 0))
    smp = int(float(row[hdr["# Samples"]] or 0))
    a = acc[kern]
    a["ops"][base][0] += ex
    a["ops"][base][1] += smp
    a["n"] += ex
    a["samples"] += smp
    for h, i in hdr.items():
        if h.startswith("stall_") and "Not Issued" not in h:
            a["stalls"][h] += int(float(row[i] or 0))
for k, a in acc.items():
    print("==", k[:110])
    print("   warp-instructions executed: %d   samples: %d" % (a["n"], a["samples"]))
    for op, (ex, smp) in sorted(a["ops"].items(), key=lambda kv: -kv[1][0])[:top]:
        print("   %-14s %12d  %5.1f%%   samples %5.1f%%" % (op, ex, 100.0 * ex / max(a["n"], 1), 100.0 * smp / max(a["samples"], 1)))
    print("   stalls:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / max(a["samples"], 1)) for h, v in sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:8]))
