#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench38_n4.log 2>&1
echo "bench n4 exit $?"; grep '^{' gpurun_out/bench38_n4.log | cut -c1-300
