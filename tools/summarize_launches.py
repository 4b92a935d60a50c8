#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: summarize_launches.py launches.csv [skip_first_n] > profiles/<name>.md"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v_us = v / 1e3 if unit in ("ns", "nsecond") else v * (1e3 if unit in ("ms", "msecond") else 1.0)
    rows.append((int(r["ID"]), r["Kernel Name"], v_us))
rows = rows[skip:]
agg = defaultdict(lambda: [0, 0.0])
for _, name, us in rows:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"<.*", "<>", short) if len(short) > 70 else short
    agg[short][0] += 1
    agg[short][1] += us
total = sum(v[1] for v in agg.values())
print(f"launches: {len(rows)}  total device time: {total/1e3:.3f} ms (cold-cache, serialised under ncu)\n")
print("| kernel | launches | total ms | share |")
print("|---|---:|---:|---:|")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {us/1e3:.3f} | {100*us/total:.1f}% |")
