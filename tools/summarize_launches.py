#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` log by kernel name: count, total, share."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
order = []
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = val * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9, "second": 1e9}.get(unit, 1)
    short = re.sub(r"<.*", "", name)[:90]
    tot[short][0] += 1
    tot[short][1] += ns
    order.append((int(r["ID"]), short, ns))
total = sum(v[1] for v in tot.values())
print("total kernels %d, total device time %.3f ms" % (sum(v[0] for v in tot.values()), total / 1e6))
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%6.2f%%  %9.3f ms  n=%5d  avg %8.1f us  %s" % (100 * t / total, t / 1e6, n, t / n / 1e3, k))
