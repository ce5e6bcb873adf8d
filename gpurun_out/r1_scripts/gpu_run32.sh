#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fast_path or graph_fusion or module_f or gluon_generic" > gpurun_out/test32.log 2>&1; echo "tests exit $?"; tail -12 gpurun_out/test32.log | cut -c1-300
timeout 300 python examples/image-classification/train_imagenet.py --gpus 0 --benchmark 1 --dtype bfloat16 --batch-size 128 --num-layers 50 --disp-batches 10 --num-epochs 1 > gpurun_out/train_imagenet32.log 2>&1
echo "train_imagenet exit $?"; grep -i "samples/sec" gpurun_out/train_imagenet32.log | tail -3
timeout 300 python examples/image-classification/train_imagenet.py --gpus 0 --benchmark 1 --dtype bfloat16 --batch-size 256 --num-layers 50 --disp-batches 10 --num-epochs 1 > gpurun_out/train_imagenet32b.log 2>&1
echo "train_imagenet b256 exit $?"; grep -i "samples/sec" gpurun_out/train_imagenet32b.log | tail -3
