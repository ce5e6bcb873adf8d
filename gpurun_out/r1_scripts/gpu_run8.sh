#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t8_all.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/t8_all.log
timeout 900 python tools/microbench.py > gpurun_out/micro8.log 2>&1; echo "micro exit $?"; grep -E "res=" gpurun_out/micro8.log
cp gpurun_out/microbench.json gpurun_out/microbench8.json
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench8.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench8.log | cut -c1-300
