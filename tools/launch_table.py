#!/usr/bin/env python
"""Per-launch table (time, tensor-pipe %, DRAM bytes) of the LAST forward step in an ncu multi-metric launch list (tools/gpu.sh launches:N).
usage: python tools/launch_table.py gpurun_out/launches_b1.csv [other.csv ...]   (several files: times side by side)"""
import collections, csv, re, sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    L = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = L.setdefault(r["ID"], {"name": re.sub(r"^void ", "", re.sub(r"\(.*", "", r["Kernel Name"])).strip(), "grid": r["Grid Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if r["Metric Name"].startswith("gpu__time"):
            v = v / 1000.0 if u.startswith("n") else (v * 1000.0 if u.startswith("m") else v)
        d[r["Metric Name"]] = v
    ids = list(L)
    starts = [i for i, k in enumerate(ids) if "conv_first" in L[k]["name"]]
    return [L[k] for k in ids[starts[-1]:]]


tabs = [load(p) for p in sys.argv[1:]]
print("| # | kernel | grid | " + " | ".join("us | tensor % | DRAM rd MB" for _ in tabs) + " |")
print("|---:|---|---|" + "---:|---:|---:|" * len(tabs))
tot = [0.0] * len(tabs)
for i, d in enumerate(tabs[0]):
    cells = []
    for j, t in enumerate(tabs):
        e = t[i] if i < len(t) and t[i]["name"] == d["name"] else None
        if e is None:
            cells.append("- | - | -")
            continue
        us = e["gpu__time_duration.sum"]
        tot[j] += us
        cells.append("%.1f | %.1f | %.1f" % (us, e.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0), e.get("dram__bytes_read.sum", 0.0) / 1e6))
    print("| %d | `%s` | %s | %s |" % (i, d["name"], d["grid"], " | ".join(cells)))
print("\ntotals (us): " + ", ".join("%.1f" % t for t in tot))
