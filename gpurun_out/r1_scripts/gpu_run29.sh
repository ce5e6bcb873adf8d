#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fast_path or graph_fusion or module_f or gluon_generic or gluon_fused" > gpurun_out/test29.log 2>&1; echo "tests exit $?"; tail -12 gpurun_out/test29.log | cut -c1-300
timeout 300 python tools/prof_module.py --steps 8 > gpurun_out/prof_module29.log 2>&1; echo "prof exit $?"; head -1 gpurun_out/prof_module29.log
timeout 300 python examples/image-classification/train_imagenet.py --gpus 0 --benchmark 1 --dtype bfloat16 --batch-size 128 --num-layers 50 --disp-batches 10 --num-epochs 1 > gpurun_out/train_imagenet29.log 2>&1
echo "train_imagenet exit $?"; grep -i "samples/sec" gpurun_out/train_imagenet29.log | tail -3
