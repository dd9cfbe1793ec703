"""Top stall sites of one kernel in an ncu source-page CSV.  Usage: python scripts/ncu_top.py <src.csv> <kernel substr> [N] [lo hi]"""
import csv
import sys

path, sub = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
lo = int(sys.argv[4], 16) if len(sys.argv) > 4 else None
hi = int(sys.argv[5], 16) if len(sys.argv) > 5 else None
hdr = None
rows = []
take = False
for row in csv.reader(open(path)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        take = sub in row[1] and not rows
        continue
    if row[0] == "Address":
        hdr = {h: i for i, h in enumerate(row)}
        continue
    if take:
        rows.append(row)
base = int(rows[0][hdr["Address"]], 16)
tot = sum(int(r[hdr["# Samples"]]) for r in rows)
print("total samples", tot, "instructions", len(rows))


def show(r):
    a = int(r[hdr["Address"]], 16) - base
    st = {h[6:]: int(r[i]) for h, i in hdr.items() if h.startswith("stall_") and "Not" not in h and int(r[i]) > 0}
    st = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print("%05x" % a, r[hdr["# Samples"]].rjust(6), "exec", r[hdr["Instructions Executed"]].rjust(9), r[hdr["Source"]].strip()[:52].ljust(52), st)


if lo is not None:
    for r in rows:
        a = int(r[hdr["Address"]], 16) - base
        if lo <= a <= hi:
            show(r)
else:
    for r in sorted(rows, key=lambda r: -int(r[hdr["# Samples"]]))[:N]:
        show(r)
