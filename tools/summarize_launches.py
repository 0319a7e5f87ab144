#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.md"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    name = re.sub(r"\(.*$", "", r["Kernel Name"])
    rows.append((name, ns))
agg = defaultdict(lambda: [0, 0.0])
for n, ns in rows:
    agg[n][0] += 1
    agg[n][1] += ns
tot = sum(v[1] for v in agg.values())
print(f"# launch list summary: {sys.argv[1]}\n")
print(f"{len(rows)} launches, {tot/1e6:.3f} ms total device time (ncu per-launch times are cold-cache and serialised: compare shares)\n")
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n}` | {c} | {ns/1e6:.3f} | {100*ns/tot:.1f}% |")
