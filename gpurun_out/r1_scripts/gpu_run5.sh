#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
for grp in "bn" "fused_allreduce or gluon_fused" "linear_sinks or gemm_nn_with"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "$grp" > "gpurun_out/t5_${name}.log" 2>&1
  echo "group [$grp] exit $?"; tail -4 "gpurun_out/t5_${name}.log"
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench5.log 2>&1; echo "bench exit $?"; tail -2 gpurun_out/bench5.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 1 -o gpurun_out/prof_gemm_fwd_802816x256x64 -f \
   python tools/prof_gemm.py gemm_fwd 802816 256 64 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 1 -o gpurun_out/prof_gemm_fwd_8192 -f \
   python tools/prof_gemm.py gemm_fwd 8192 8192 8192 > gpurun_out/ncu_gemm2.log 2>&1; echo "ncu gemm2 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bwd_reduce_kernel -s 1 -c 1 -o gpurun_out/prof_bn_bwd_reduce -f \
   python tools/prof_gemm.py bn > gpurun_out/ncu_bn.log 2>&1; echo "ncu bn exit $?"
ls -la gpurun_out/*.ncu-rep
