#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
for grp in "gemm or linear" "bn" "maxpool or softmax or sgd_kernel or adam or elementwise" "fused_allreduce or gluon_fused"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "$grp" > "gpurun_out/t4_${name}.log" 2>&1
  echo "group [$grp] exit $?"; tail -4 "gpurun_out/t4_${name}.log"
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench4.log 2>&1; echo "bench exit $?"; tail -3 gpurun_out/bench4.log
MXB200_STEM_PAD=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e > gpurun_out/bench4_nopad.log 2>&1; echo "bench nopad exit $?"; tail -2 gpurun_out/bench4_nopad.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches4.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_bench4.log 2>&1; echo "ncu exit $?"
python tools/summarize_launches.py gpurun_out/launches4.csv > gpurun_out/launch_summary4.txt 2>&1; head -40 gpurun_out/launch_summary4.txt
timeout 600 python tools/microbench.py > gpurun_out/micro4.log 2>&1; echo "micro exit $?"; grep -E "fwd|dgrad|wgrad|ours" gpurun_out/micro4.log | tail -50
