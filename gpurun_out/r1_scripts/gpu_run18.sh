#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bert18.csv python tools/bench_bert.py --steps 2 --warmup 3 > gpurun_out/ncu_bert18.log 2>&1; echo "ncu bert exit $?"
python tools/summarize_launches.py gpurun_out/launches_bert18.csv > gpurun_out/launch_summary_bert18.txt 2>&1; head -24 gpurun_out/launch_summary_bert18.txt
