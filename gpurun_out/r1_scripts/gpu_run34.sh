#!/bin/bash
# final 1-GPU validation of the round: full GPU test suite, smoke, headline bench, ncu captures of the CTA-pair GEMM
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/test34.log 2>&1; echo "pytest gpu exit $?"; tail -4 gpurun_out/test34.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke34.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke34.log | cut -c1-200
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench34.log 2>&1; echo "bench exit $?"; grep '^{' gpurun_out/bench34.log | cut -c1-400
for spec in "gemm_fwd 8192 8192 8192" "gemm_wgrad 12544 2048 512"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 1 -o gpurun_out/prof34_${1}_${2}x${3}x${4} -f python tools/prof_gemm.py $1 $2 $3 $4 > gpurun_out/ncu34_$1_$2.log 2>&1; echo "ncu $spec exit $?"
done
ls -la gpurun_out/prof34_*.ncu-rep
