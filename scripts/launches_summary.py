#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.

    python scripts/launches_summary.py profiles/r01f_launches.csv > profiles/r01f_launches_summary.txt
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[start + 2:]:
    if len(r) > iv:
        try:
            agg.setdefault(r[ik], []).append(float(r[iv].replace(",", "")) / 1e3)
        except ValueError:
            pass
tot = sum(sum(v) for v in agg.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none -c 500: python bench.py --steps 2 --warmup 1 --no-cpu-baseline --large-batch 0")
print("# (cold-cache, serialised launches: compare SHARES with bench.py's event-timed 'kernels', not absolutes)")
print(f"{'kernel':80s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:78]:80s} {len(v):8d} {sum(v):12.1f} {sum(v) / len(v):10.1f} {sum(v) / tot:7.3f}")
