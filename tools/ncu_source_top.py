#!/usr/bin/env python
"""Top stall sites of one kernel from an `ncu --page source --csv` export (SASS view).
usage: python tools/ncu_source_top.py file.csv <kernel index> [top N] [context lines]"""
import csv, sys
csv.field_size_limit(10**9)
path, kidx = sys.argv[1], int(sys.argv[2])
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
ctx = int(sys.argv[4]) if len(sys.argv) > 4 else 0
kernels, cur, hdr = [], None, None
for row in csv.reader(open(path)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        kernels.append(cur)
        hdr = None
    elif row[0] == "Address":
        hdr = row
        cur["hdr"] = hdr
    elif cur is not None and hdr is not None:
        cur["rows"].append(row)
k = kernels[kidx]
h = {n: i for i, n in enumerate(k["hdr"])}
rows = k["rows"]
tot = sum(int(r[h["# Samples"]] or 0) for r in rows)
print(k["name"][:100], "instructions:", len(rows), "samples:", tot)
stalls = [n for n in k["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
order = sorted(range(len(rows)), key=lambda i: -int(rows[i][h["# Samples"]] or 0))[:topn]
for i in sorted(order):
    r = rows[i]
    s = int(r[h["# Samples"]] or 0)
    top = sorted(((int(r[h[n]] or 0), n) for n in stalls), reverse=True)[:3]
    for j in range(max(0, i - ctx), i):
        print("      %5d        %s" % (j, rows[j][h["Source"]].strip()[:90]))
    print("%5.1f%% %5d exec=%-7s %-70s %s" % (100.0 * s / max(tot, 1), i, r[h["Instructions Executed"]], r[h["Source"]].strip()[:70],
                                        " ".join("%s=%d" % (n[6:], v) for v, n in top if v)))
