#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches23.csv python bench.py --steps 1 --warmup 3 --no-e2e > gpurun_out/ncu_bench23.log 2>&1; echo "ncu bench exit $?"
python tools/summarize_launches.py gpurun_out/launches23.csv > gpurun_out/launch_summary23.txt 2>&1; head -44 gpurun_out/launch_summary23.txt
