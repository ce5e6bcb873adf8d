#!/bin/bash
# 4-GPU: elastic run through the gluon API (kvstore 'elastic_device')
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
R=$(pwd)
rm -rf /tmp/elasticg && mkdir -p /tmp/elasticg && cd /tmp/elasticg
timeout 420 python $R/tools/launch.py -n 2 -s 1 --launcher local --elastic-training-enabled True python $R/tools/elastic_gluon_worker.py /tmp/elasticg 2,4,4,2,2 6 64 > $R/gpurun_out/elastic_gluon4.log 2>&1; echo "elastic gluon exit $?"
cd $R
tail -6 gpurun_out/elastic_gluon4.log | cut -c1-300
mkdir -p gpurun_out/elastic_gluon4; cp /tmp/elasticg/out_*.json /tmp/elasticg/host_worker_log gpurun_out/elastic_gluon4/ 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/elastic_gluon4/out_*.json")):
    d=json.load(open(f)); print(d["tag"], "new" if d["new"] else "base", [(e["epoch"],e["rank"],e["world"],round(e["loss_first"],3),round(e["loss_last"],3),e["weights_identical"]) for e in d["epochs"]])
print(open("gpurun_out/elastic_gluon4/host_worker_log").read())
PY
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "fused_resnet_symbol" > gpurun_out/t21.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/t21.log
