#!/usr/bin/env python
"""Top stall sites of an .ncu-rep source page (SASS view).  usage: ncu_source_top.py report.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:N]
for i in sorted(order):
    r = body[i]
    s = int(r[ix["# Samples"]] or 0)
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print("%5d %5.1f%%  exec=%-8s %-70s %s" % (i, 100.0 * s / max(tot, 1), r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:70],
                                       " ".join("%s:%d" % (n, v) for v, n in st if v)))
