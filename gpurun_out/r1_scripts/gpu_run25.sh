#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"maxpool|subsample|im2col" -c 40 --csv --log-file gpurun_out/launches25.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_bench25.log 2>&1; echo "ncu exit $?"
python tools/summarize_launches.py gpurun_out/launches25.csv | head -8
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench25.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench25.log | cut -c1-300
timeout 600 python bench.py --steps 30 --warmup 5 --no-e2e > gpurun_out/bench25b.log 2>&1; echo "bench again exit $?"; tail -1 gpurun_out/bench25b.log | cut -c1-200
