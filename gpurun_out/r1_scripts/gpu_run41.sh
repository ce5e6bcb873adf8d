#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/test41.log 2>&1; echo "pytest gpu exit $?"; tail -3 gpurun_out/test41.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke41.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke41.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench41.log 2>&1; echo "bench exit $?"; grep '^{' gpurun_out/bench41.log | cut -c1-260
