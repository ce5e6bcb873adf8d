#!/bin/bash
# 2 GPUs: reference-style script under torchrun (auto process group + fused Module.update across ranks), then the headline bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 examples/image-classification/train_imagenet.py --gpus 0,1 --benchmark 1 --dtype bfloat16 --batch-size 256 --num-layers 50 --disp-batches 10 --num-epochs 1 --kv-store device > gpurun_out/train_imagenet35_n2.log 2>&1
echo "train_imagenet n2 exit $?"; grep -i "samples/sec" gpurun_out/train_imagenet35_n2.log | tail -4; grep -i "error\|Traceback" gpurun_out/train_imagenet35_n2.log | head -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench35_n2.log 2>&1
echo "bench n2 exit $?"; grep '^{' gpurun_out/bench35_n2.log | cut -c1-300
