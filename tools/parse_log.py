#!/usr/bin/env python
"""Summarise a training log written by Module.fit / Speedometer (reference: tools/parse_log.py): per epoch
train / validation metric, time cost and mean speed, as a markdown table."""
import re
import sys

if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit("usage: parse_log.py logfile [metric=accuracy]")
    metric = sys.argv[2] if len(sys.argv) > 2 else "accuracy"
    rows = {}
    pats = {"train": re.compile(r"Epoch\[(\d+)\] Train-%s=([\d.eE+-]+)" % re.escape(metric)),
            "val": re.compile(r"Epoch\[(\d+)\] Validation-%s=([\d.eE+-]+)" % re.escape(metric)),
            "time": re.compile(r"Epoch\[(\d+)\] Time cost=([\d.]+)"),
            "speed": re.compile(r"Epoch\[(\d+)\] Batch \[\d+\]\s+Speed: ([\d.]+) samples/sec")}
    for line in open(sys.argv[1]):
        for k, p in pats.items():
            m = p.search(line)
            if m:
                r = rows.setdefault(int(m.group(1)), {"speed": []})
                if k == "speed":
                    r["speed"].append(float(m.group(2)))
                else:
                    r[k] = float(m.group(2))
    print("| epoch | train-%s | val-%s | time (s) | speed (samples/s) |" % (metric, metric))
    print("| --- | --- | --- | --- | --- |")
    for e in sorted(rows):
        r = rows[e]
        sp = sum(r["speed"]) / len(r["speed"]) if r["speed"] else float("nan")
        print("| %d | %s | %s | %s | %.1f |" % (e, r.get("train", ""), r.get("val", ""), r.get("time", ""), sp))
