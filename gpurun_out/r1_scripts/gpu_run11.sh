#!/bin/bash
# 4-GPU: elastic 2->4->4->2->2 on GPUs, fused allreduce check at N=4, bench N=4, kvstore bandwidth tool
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
R=$(pwd)
rm -rf /tmp/elastic && mkdir -p /tmp/elastic && cd /tmp/elastic
timeout 420 python $R/tools/launch.py -n 2 -s 1 --launcher local --elastic-training-enabled True python $R/tools/elastic_gpu_worker.py /tmp/elastic 2,4,4,2,2 6 64 > $R/gpurun_out/elastic4.log 2>&1; echo "elastic exit $?"
cd $R
tail -15 gpurun_out/elastic4.log | cut -c1-300
mkdir -p gpurun_out/elastic4; cp /tmp/elastic/out_*.json /tmp/elastic/host_worker_log gpurun_out/elastic4/ 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/elastic4/out_*.json")):
    d=json.load(open(f)); print(d["tag"], "new" if d["new"] else "base", [(e["epoch"],e["rank"],e["world"],round(e["loss_first"],3),round(e["loss_last"],3),e["weights_identical"]) for e in d["epochs"]])
print(open("gpurun_out/elastic4/host_worker_log").read())
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 tools/check_multigpu.py > gpurun_out/multigpu4.log 2>&1; echo "multigpu check exit $?"; grep '^{' gpurun_out/multigpu4.log | cut -c1-420
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench11_n4.log 2>&1; echo "bench n4 exit $?"; grep '^{' gpurun_out/bench11_n4.log | cut -c1-330
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 tools/bandwidth/measure.py --network resnet --num-layers 152 --mode fused --num-batches 5 --json gpurun_out/bandwidth_fused_n4.json > gpurun_out/bw_fused4.log 2>&1; echo "bw fused exit $?"; grep "iter" gpurun_out/bw_fused4.log | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 tools/bandwidth/measure.py --network resnet --num-layers 152 --mode kv --num-batches 3 --json gpurun_out/bandwidth_kv_n4.json > gpurun_out/bw_kv4.log 2>&1; echo "bw kv exit $?"; grep "iter" gpurun_out/bw_kv4.log | tail -2
timeout 300 python tools/bandwidth/measure.py --network resnet --num-layers 152 --gpus 0,1,2,3 --mode kv --num-batches 3 --json gpurun_out/bandwidth_kv_singleproc_n4.json > gpurun_out/bw_kv4sp.log 2>&1; echo "bw kv single-process exit $?"; grep "iter" gpurun_out/bw_kv4sp.log | tail -2
