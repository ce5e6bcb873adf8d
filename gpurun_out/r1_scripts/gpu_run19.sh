#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t19_all.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/t19_all.log
timeout 300 python tools/bench_bert.py --steps 20 --warmup 5 > gpurun_out/bert19.log 2>&1; echo "bert exit $?"; tail -1 gpurun_out/bert19.log | cut -c1-300
timeout 300 python tools/bench_bert.py --steps 20 --warmup 5 --batch 32 --seq 512 --masked 76 > gpurun_out/bert19_s512.log 2>&1; echo "bert seq512 exit $?"; tail -1 gpurun_out/bert19_s512.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench19.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench19.log | cut -c1-300
