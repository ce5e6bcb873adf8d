#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t13_all.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/t13_all.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench13.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench13.log | cut -c1-300
timeout 900 python tools/microbench.py > gpurun_out/micro13.log 2>&1; echo "micro exit $?"; grep -E "res=" gpurun_out/micro13.log
cp gpurun_out/microbench.json gpurun_out/microbench13.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches13.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_bench13.log 2>&1; echo "ncu bench exit $?"
python tools/summarize_launches.py gpurun_out/launches13.csv > gpurun_out/launch_summary13.txt 2>&1; head -14 gpurun_out/launch_summary13.txt
timeout 600 python examples/image-classification/benchmark_score.py --networks resnet50_v1b,resnet50_v1,resnet152_v1 --batch-sizes 32,128 --dtype bfloat16 --json gpurun_out/benchmark_score13.json > gpurun_out/score13.log 2>&1; echo "score exit $?"; grep "images/sec" gpurun_out/score13.log
