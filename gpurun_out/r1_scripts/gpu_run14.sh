#!/bin/bash
# 8-GPU: bench N=8, fused allreduce check at N=8, elastic 4->8->4, kvstore bandwidth tool (fused + kv)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
R=$(pwd)
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "stem or subsample or prefetcher" > gpurun_out/t14_stem.log 2>&1; echo "pytest stem exit $?"; tail -3 gpurun_out/t14_stem.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench14_n8.log 2>&1; echo "bench n8 exit $?"; grep '^{' gpurun_out/bench14_n8.log | cut -c1-330
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 tools/check_multigpu.py > gpurun_out/multigpu8.log 2>&1; echo "multigpu check exit $?"; grep '^{' gpurun_out/multigpu8.log | cut -c1-420
rm -rf /tmp/elastic8 && mkdir -p /tmp/elastic8 && cd /tmp/elastic8
timeout 400 python $R/tools/launch.py -n 4 -s 1 --launcher local --elastic-training-enabled True python $R/tools/elastic_gpu_worker.py /tmp/elastic8 4,8,8,4,4 6 128 > $R/gpurun_out/elastic8.log 2>&1; echo "elastic exit $?"
cd $R
mkdir -p gpurun_out/elastic8; cp /tmp/elastic8/out_*.json /tmp/elastic8/host_worker_log gpurun_out/elastic8/ 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/elastic8/out_*.json")):
    d=json.load(open(f)); print(d["tag"], "new" if d["new"] else "base", [(e["epoch"],e["rank"],e["world"],round(e["loss_last"],3),e["weights_identical"]) for e in d["epochs"]])
print(open("gpurun_out/elastic8/host_worker_log").read())
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/bandwidth/measure.py --network resnet --num-layers 200 --mode fused --num-batches 5 --json gpurun_out/bandwidth_fused_n8.json > gpurun_out/bw_fused8.log 2>&1; echo "bw fused exit $?"; grep "iter" gpurun_out/bw_fused8.log | tail -2
timeout 300 python tools/bandwidth/measure.py --network resnet --num-layers 200 --gpus 0,1,2,3,4,5,6,7 --mode kv --num-batches 3 --json gpurun_out/bandwidth_kv_singleproc_n8.json > gpurun_out/bw_kv8sp.log 2>&1; echo "bw kv single-process exit $?"; grep "iter" gpurun_out/bw_kv8sp.log | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 tools/bench_bert.py --steps 10 --warmup 3 > gpurun_out/bert14_n8.log 2>&1; echo "bert n8 exit $?"; grep '^{' gpurun_out/bert14_n8.log | cut -c1-400
