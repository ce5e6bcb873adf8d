#!/bin/bash
# new fast paths: numerics, then the generic NCHW symbolic ResNet-50 (Module.fit) with and without them
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "fast_paths or graph_fusion or module_fit or gluon_generic" > gpurun_out/test27.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/test27.log
for mode in 1 0; do
  MXB200_OP_FASTPATH=$mode MXB200_GRAPH_FUSION=$mode timeout 300 python examples/image-classification/train_imagenet.py --gpus 0 --benchmark 1 --dtype bfloat16 --batch-size 128 --num-layers 50 --disp-batches 10 --num-epochs 1 > gpurun_out/train_imagenet27_$mode.log 2>&1
  echo "train_imagenet fast=$mode exit $?"; grep -i "samples/sec" gpurun_out/train_imagenet27_$mode.log | tail -3; tail -2 gpurun_out/train_imagenet27_$mode.log | cut -c1-300
done
