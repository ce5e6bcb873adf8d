#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the few numbers we track in profiles/.

usage: python tools/ncu_summary.py report.ncu-rep [substr ...]
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared",
        "smsp__average_warps_issue_stalled", "smsp__issue_active.avg.pct", "lts__throughput.avg.pct",
        "l1tex__throughput.avg.pct", "sm__cycles_active.avg", "smsp__inst_executed.sum ",
        "smsp__cycles_active.avg ", "sm__cycles_elapsed.max"]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("## kernel:", dict(zip(hdr, vals)).get("Kernel Name", "?")[:100])
        for h, u, v in zip(hdr, units, vals):
            hs = h + " "
            if any(w in hs for w in WANT + extra):
                print("  %-90s %-14s %s" % (h, u, v))


if __name__ == "__main__":
    main()
