#!/usr/bin/env python
"""Per-layer bandwidth of the BatchNorm kernels inside one eager ResNet-50 training step, from an ncu launch log
(`ncu --metrics gpu__time_duration.sum,launch__grid_size --csv ... python bench.py --no-graph`): every forward apply,
backward reduce and backward apply is matched to its layer (order of the functional model) and its minimum traffic.

    python tools/step_roofline.py launches.csv [batch]"""
import csv
import re
import sys
from collections import defaultdict

B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = defaultdict(dict)
for r in csv.DictReader(lines):
    rid = int(r["ID"])
    rows[rid]["name"] = re.sub(r"\(.*", "", r["Kernel Name"])
    if r["Metric Name"] == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        rows[rid]["us"] = v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(r.get("Metric Unit", "ns"), 1e-3)
ids = sorted(rows)
ends = [i for i in ids if "ar::fused_kernel<float, __nv_bfloat16" in rows[i]["name"]]
a, b = ends[-2], ends[-1]
seq = [(rows[i]["name"], rows[i]["us"]) for i in ids if a < i <= b]
print("launches in the step: %d, device time %.2f ms" % (len(seq), sum(s[1] for s in seq) / 1e3))

layers = [(B * 112 * 112, 64, 0, 1)]                      # (rows, channels, residual, relu)


def stage(n, hw_in, hw, mid, out):
    L = []
    for blk in range(n):
        h1 = hw_in if blk == 0 else hw
        L.append((B * h1 * h1, mid, 0, 1))                 # bn1 (1x1 conv at the input resolution; v1b strides the 3x3)
        if blk == 0:
            L.append((B * hw * hw, out, 0, 0))             # down-sample BatchNorm
        L.append((B * hw * hw, mid, 0, 1))                 # bn2
        L.append((B * hw * hw, out, 1, 1))                 # bn3 + residual + ReLU
    return L


layers += stage(3, 56, 56, 64, 256) + stage(4, 56, 28, 128, 512) + stage(6, 28, 14, 256, 1024) + stage(3, 14, 7, 512, 2048)
fwd = [s for s in seq if "elem_bulk_kernel<0" in s[0]]
bwd_apply = [s for s in seq if "elem_bulk_kernel<1" in s[0]]
bwd_red = [s for s in seq if "bwd_reduce_bulk_kernel" in s[0]]
print("BatchNorm layers %d; forward applies %d, backward reduces %d, backward applies %d" % (len(layers), len(fwd), len(bwd_red), len(bwd_apply)))
tot = {"fwd": [0, 0], "red": [0, 0], "app": [0, 0]}
print("%9s %5s %3s | %18s | %18s | %18s" % ("rows", "C", "res", "fwd apply", "bwd reduce", "bwd apply"))
for L, f, r_, ap in zip(layers, fwd, reversed(bwd_red), reversed(bwd_apply)):
    M, C, res, relu = L
    e = M * C * 2
    bf = e * (2 + res)                  # read x (+ residual), write y
    br = e * 2                          # read dy, x
    ba = e * (3 + res)                  # read dy, x; write dx (+ d residual)
    for k, byt, us in (("fwd", bf, f[1]), ("red", br, r_[1]), ("app", ba, ap[1])):
        tot[k][0] += byt
        tot[k][1] += us
    print("%9d %5d %3d | %7.1f us %5.2f TB/s | %7.1f us %5.2f TB/s | %7.1f us %5.2f TB/s" % (
        M, C, res, f[1], bf / f[1] / 1e6, r_[1], br / r_[1] / 1e6, ap[1], ba / ap[1] / 1e6))
for k, name in (("fwd", "forward apply"), ("red", "backward reduce"), ("app", "backward apply")):
    print("%-16s %6.2f ms per step, %.2f TB/s aggregate (minimum traffic / time)" % (name, tot[k][1] / 1e3, tot[k][0] / tot[k][1] / 1e6))
