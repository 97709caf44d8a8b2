#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into a per-kernel table
for ONE forward step (between two consecutive conv_first_kernel launches).  Per-launch times under ncu are
cold-cache and serialised: compare SHARES, not absolutes.

usage: python tools/summarize_launches.py gpurun_out/launches_rNN.csv > profiles/launches_rNN_summary.md
"""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    out = []
    for r in rows:
        t = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if u in ("ns", "nsecond"):
            t /= 1000.0
        elif u in ("ms", "msecond"):
            t *= 1000.0
        elif u in ("s", "second"):
            t *= 1e6
        out.append((re.sub(r"\(.*", "", r["Kernel Name"]).strip(), r.get("Grid Size", ""), r.get("Block Size", ""), t))
    return out


def main():
    rows = load(sys.argv[1])
    starts = [i for i, r in enumerate(rows) if "conv_first" in r[0]]
    if len(starts) < 2:
        print("need at least two steps in the capture", file=sys.stderr)
        sys.exit(1)
    s, e = starts[-2], starts[-1]
    step = rows[s:e]
    total = sum(r[3] for r in step)
    agg = collections.OrderedDict()
    for name, grid, block, t in step:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    print("# Launch list of one forward step (%s)\n" % sys.argv[1])
    print("%d launches, %.1f us summed kernel time (ncu-serialised, cold caches; shares are what matters)\n" % (len(step), total))
    print("| kernel | launches | total us | share |")
    print("|---|---:|---:|---:|")
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% |" % (name, c, t, 100.0 * t / total))
    print("\n## In launch order\n")
    print("| # | kernel | grid | us |")
    print("|---:|---|---|---:|")
    for i, (name, grid, block, t) in enumerate(step):
        print("| %d | `%s` | %s | %.1f |" % (i, name, grid, t))


if __name__ == "__main__":
    main()
