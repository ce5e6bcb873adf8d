#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t17_all.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/t17_all.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench17.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench17.log | cut -c1-1900
timeout 600 python bench.py --steps 20 --warmup 5 --batch 256 --no-e2e > gpurun_out/bench17_b256.log 2>&1; echo "bench b256 exit $?"; tail -1 gpurun_out/bench17_b256.log | cut -c1-250
timeout 300 python tools/bench_bert.py --steps 10 --warmup 3 > gpurun_out/bert17.log 2>&1; echo "bert exit $?"; tail -1 gpurun_out/bert17.log | cut -c1-300
timeout 900 python tools/microbench.py > gpurun_out/micro17.log 2>&1; echo "micro exit $?"; grep -E "res=" gpurun_out/micro17.log
cp gpurun_out/microbench.json gpurun_out/microbench17.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
