#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t12_all.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/t12_all.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench12.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench12.log | cut -c1-1800
timeout 900 python tools/microbench.py > gpurun_out/micro12.log 2>&1; echo "micro exit $?"; grep -E "res=" gpurun_out/micro12.log
cp gpurun_out/microbench.json gpurun_out/microbench12.json
timeout 300 python tools/bench_bert.py --steps 10 --warmup 3 > gpurun_out/bert12.log 2>&1; echo "bert exit $?"; tail -2 gpurun_out/bert12.log | cut -c1-900
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches12.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_bench12.log 2>&1; echo "ncu bench exit $?"
python tools/summarize_launches.py gpurun_out/launches12.csv > gpurun_out/launch_summary12.txt 2>&1; head -34 gpurun_out/launch_summary12.txt
