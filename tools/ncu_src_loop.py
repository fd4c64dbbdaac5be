#!/usr/bin/env python
"""Per-instruction stall samples of one kernel from `ncu --page source --csv --print-source sass` output.
usage: ncu_src_loop.py file.csv [kernel-index] [first-line last-line]   (prints the instructions with samples)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
# split into kernels
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; kernels.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
ki = int(sys.argv[2]) if len(sys.argv) > 2 else 0
k = kernels[ki]
h = k["hdr"]
col = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
print(k["name"], len(k["rows"]), "instructions; kernels:", len(kernels))
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else len(k["rows"])
tot = sum(int(r[col["# Samples"]]) for r in k["rows"])
print("total samples", tot)
for n, r in enumerate(k["rows"][lo:hi], lo):
    s = int(r[col["# Samples"]])
    top = sorted(((int(r[col[x]]), x[6:]) for x in stalls), reverse=True)[:3]
    tops = " ".join("%s=%d" % (b, a) for a, b in top if a)
    print("%5d %6d x%-6s %-60s %s" % (n, s, r[col["Instructions Executed"]], r[col["Source"]].strip()[:60], tops))
