#!/bin/bash
# 8-GPU final: bench N=8 at the default batch (512/GPU = global 4096), elastic 4->8->4 through the gluon API
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
R=$(pwd)
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench22_n8.log 2>&1; echo "bench n8 exit $?"; grep '^{' gpurun_out/bench22_n8.log | cut -c1-2000
rm -rf /tmp/elasticg8 && mkdir -p /tmp/elasticg8 && cd /tmp/elasticg8
timeout 400 python $R/tools/launch.py -n 4 -s 1 --launcher local --elastic-training-enabled True python $R/tools/elastic_gluon_worker.py /tmp/elasticg8 4,8,8,4,4 6 128 > $R/gpurun_out/elastic_gluon8.log 2>&1; echo "elastic gluon exit $?"
cd $R
mkdir -p gpurun_out/elastic_gluon8; cp /tmp/elasticg8/out_*.json /tmp/elasticg8/host_worker_log gpurun_out/elastic_gluon8/ 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/elastic_gluon8/out_*.json")):
    d=json.load(open(f)); print(d["tag"], "new" if d["new"] else "base", [(e["epoch"],e["rank"],e["world"],round(e["loss_last"],3),e["weights_identical"]) for e in d["epochs"]])
print(open("gpurun_out/elastic_gluon8/host_worker_log").read())
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 tools/check_multigpu.py > gpurun_out/multigpu8b.log 2>&1; echo "multigpu check exit $?"; grep '^{' gpurun_out/multigpu8b.log | tail -1 | cut -c1-600
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29553 tools/bench_bert.py --steps 10 --warmup 3 > gpurun_out/bert22_n8.log 2>&1; echo "bert n8 exit $?"; grep '^{' gpurun_out/bert22_n8.log | cut -c1-300
