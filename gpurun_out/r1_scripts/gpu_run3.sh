#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
# per-kernel device time of 2 steady-state steps (skip the warm-up launches)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "ncu exit $?"
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt 2>&1; head -60 gpurun_out/launch_summary.txt
