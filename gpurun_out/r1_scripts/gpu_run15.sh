#!/bin/bash
# 1-GPU: BERT kernel breakdown, final ncu --set full captures of the hot kernels, batch-size sensitivity of the bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bert15.csv python tools/bench_bert.py --steps 2 --warmup 3 > gpurun_out/ncu_bert15.log 2>&1; echo "ncu bert exit $?"
python tools/summarize_launches.py gpurun_out/launches_bert15.csv > gpurun_out/launch_summary_bert15.txt 2>&1; head -40 gpurun_out/launch_summary_bert15.txt
for spec in "gemm_fwd 802816 256 64" "gemm_fwd 8192 8192 8192" "gemm_wgrad 12544 2048 512"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 1 -o gpurun_out/prof15_${1}_${2}x${3}x${4} -f python tools/prof_gemm.py $1 $2 $3 $4 > gpurun_out/ncu15_$1_$2.log 2>&1; echo "ncu $spec exit $?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bwd_reduce_kernel -s 1 -c 1 -o gpurun_out/prof15_bn_bwd_reduce_50176x1024 -f python tools/prof_gemm.py bn 50176 1024 > gpurun_out/ncu15_bn1.log 2>&1; echo "ncu bn reduce exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bwd_apply_kernel -s 1 -c 1 -o gpurun_out/prof15_bn_bwd_apply_802816x256 -f python tools/prof_gemm.py bn 802816 256 > gpurun_out/ncu15_bn2.log 2>&1; echo "ncu bn apply exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 1 -c 1 -o gpurun_out/prof15_fused_allreduce_1gpu -f python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu15_ar.log 2>&1; echo "ncu fused exit $?"
for b in 384 512; do
  timeout 600 python bench.py --steps 20 --warmup 5 --batch $b --no-e2e > gpurun_out/bench15_b$b.log 2>&1; echo "bench batch $b exit $?"; grep '^{' gpurun_out/bench15_b$b.log | cut -c1-200
done
ls -la gpurun_out/prof15_*.ncu-rep
