#!/bin/bash
# 1-GPU: full gpu test-suite, microbench (incl. tile-width sweep), bench, fused-trainer diagnostic
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t6_all.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/t6_all.log
timeout 600 python tools/diag_fused.py 3 > gpurun_out/diag_fused.log 2>&1; echo "diag exit $?"; tail -25 gpurun_out/diag_fused.log
timeout 900 python tools/microbench.py > gpurun_out/micro6.log 2>&1; echo "micro exit $?"; grep -E "ratio|speedup|fwd_bn" gpurun_out/micro6.log | head -70
cp gpurun_out/microbench.json gpurun_out/microbench6.json
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench6.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench6.log
