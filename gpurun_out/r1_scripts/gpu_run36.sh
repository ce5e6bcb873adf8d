#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/test36.log 2>&1; echo "pytest gpu exit $?"; tail -6 gpurun_out/test36.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke36.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke36.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench36.log 2>&1; echo "bench exit $?"; grep '^{' gpurun_out/bench36.log | cut -c1-300
timeout 300 python examples/image-classification/train_imagenet.py --gpus 0 --benchmark 1 --dtype bfloat16 --batch-size 256 --num-layers 50 --disp-batches 10 --num-epochs 1 > gpurun_out/train_imagenet36.log 2>&1
echo "train_imagenet exit $?"; grep -i "samples/sec" gpurun_out/train_imagenet36.log | tail -2
