#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t7_all.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/t7_all.log
timeout 900 python tools/microbench.py > gpurun_out/micro7.log 2>&1; echo "micro exit $?"; grep -E "ratio|speedup|fwd_bn" gpurun_out/micro7.log | head -70
cp gpurun_out/microbench.json gpurun_out/microbench7.json
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench7.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench7.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches7.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench7.log 2>&1; echo "ncu bench exit $?"
python tools/summarize_launches.py gpurun_out/launches7.csv > gpurun_out/launch_summary7.txt 2>&1; head -40 gpurun_out/launch_summary7.txt
