#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t9_all.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/t9_all.log
for shp in "12544 2048" "50176 1024" "200704 512" "802816 64"; do
  set -- $shp
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bn_${1}_${2}.csv python tools/prof_gemm.py bn $1 $2 > /dev/null 2>&1
  echo "== BN M=$1 C=$2"; python tools/summarize_launches.py gpurun_out/bn_${1}_${2}.csv | grep -E "bn::|at::" | head -8
done
timeout 900 python tools/microbench.py > gpurun_out/micro9.log 2>&1; echo "micro exit $?"; grep -E "res=" gpurun_out/micro9.log
cp gpurun_out/microbench.json gpurun_out/microbench9.json
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench9.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench9.log | cut -c1-300
MXB200_STEM=cudnn timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench9_cudnnstem.log 2>&1; echo "bench(cudnn stem) exit $?"; tail -1 gpurun_out/bench9_cudnnstem.log | cut -c1-200
