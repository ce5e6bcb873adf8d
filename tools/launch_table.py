#!/usr/bin/env python
"""Per-launch table from an `ncu --metrics gpu__time_duration.sum,launch__grid_size --csv` log: groups launches by
(kernel, grid size) so the same kernel at different problem sizes shows up separately (which layers are slow).

    python tools/launch_table.py launches.csv [substring]"""
import csv
import re
import sys
from collections import defaultdict

with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rows = defaultdict(dict)
for r in csv.DictReader(lines):
    rid = int(r["ID"])
    rows[rid]["name"] = re.sub(r"\(.*", "", r["Kernel Name"])[:110]
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Name"] == "gpu__time_duration.sum":
        unit = r.get("Metric Unit", "ns")
        rows[rid]["ns"] = v * {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
    elif r["Metric Name"] == "launch__grid_size":
        rows[rid]["grid"] = int(v)
sub = sys.argv[2] if len(sys.argv) > 2 else ""
agg = defaultdict(lambda: [0, 0.0])
for rid, d in rows.items():
    if "ns" not in d or sub not in d["name"]:
        continue
    k = (d["name"], d.get("grid", 0))
    agg[k][0] += 1
    agg[k][1] += d["ns"]
tot = sum(v[1] for v in agg.values())
print("launches %d, device time %.3f ms" % (sum(v[0] for v in agg.values()), tot / 1e6))
for (name, grid), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
    print("%6.2f%% %9.3f ms n=%4d avg %8.1f us grid %6d  %s" % (100 * t / tot, t / 1e6, n, t / n / 1e3, grid, name))
