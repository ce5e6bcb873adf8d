#!/bin/bash
# 2-GPU: fused P2P allreduce+update correctness/timing, bench at N=2 (torchrun), BN + new tests on rank 0 first
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t10_all.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/t10_all.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multigpu.py > gpurun_out/multigpu2.log 2>&1; echo "multigpu check exit $?"; tail -8 gpurun_out/multigpu2.log | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench10_n2.log 2>&1; echo "bench n2 exit $?"; tail -3 gpurun_out/bench10_n2.log | cut -c1-1200
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench10_n1.log 2>&1; echo "bench n1 exit $?"; tail -1 gpurun_out/bench10_n1.log | cut -c1-300
