#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "linear_autograd" > gpurun_out/t_linear.log 2>&1; echo "linear exit $?"; tail -3 gpurun_out/t_linear.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.log 2>&1; echo "bench exit $?"; tail -5 gpurun_out/bench1.log
timeout 600 python tools/microbench.py > gpurun_out/micro.log 2>&1; echo "micro exit $?"; tail -40 gpurun_out/micro.log
