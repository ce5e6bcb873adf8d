set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/c5
timeout 900 python -m pytest tests/test_multigpu.py -q > gpurun_out/c5/pytest_multigpu.txt 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/c5/pytest_multigpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 tools/bandwidth/sweep.py --max-bytes 67108864 --out gpurun_out/c5/bandwidth_sweep_n2.json > gpurun_out/c5/sweep_n2.txt 2>&1; echo "sweep rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c5/bandwidth_sweep_n2.json"))
    for r in d["rows"]:
        print(r["bytes"], {k: round(v, 1) for k, v in r.items() if k.endswith("_us")})
except Exception as e:
    print("no sweep json", e)
PY
tail -5 gpurun_out/c5/sweep_n2.txt
timeout 600 python tools/bandwidth/measure.py --gpus 0,1 --network resnet --num-layers 200 --num-batches 6 --mode kv --json gpurun_out/c5/bandwidth_kv_singleproc_n2.json > gpurun_out/c5/measure_kv_n2.txt 2>&1; echo "measure rc=$?"
tail -12 gpurun_out/c5/measure_kv_n2.txt
mkdir -p gpurun_out/c5/elastic && cd gpurun_out/c5/elastic && rm -f host_worker* && \
  timeout 900 python $GRAFT_REPO_ROOT/tools/launch.py -n 1 -s 1 --launcher local --elastic-training-enabled True \
    python $GRAFT_REPO_ROOT/tools/elastic_gluon_worker.py $GRAFT_REPO_ROOT/gpurun_out/c5/elastic --plan 1,2,1 --layers 50 --image 224 --classes 1000 --batch-per-gpu 256 --steps 10 --lr 0.05 > launch.log 2>&1; echo "elastic rc=$?"
cd $GRAFT_REPO_ROOT; tail -20 gpurun_out/c5/elastic/launch.log; python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c5/elastic/out_*.json")):
    d = json.load(open(f))
    for e in d["epochs"]:
        print(d["tag"], {k: (round(v, 3) if isinstance(v, float) else v) for k, v in e.items() if k not in ("reinit", "lr_change")})
        if "reinit" in e: print("   reinit", {k: round(v, 1) if isinstance(v, float) else v for k, v in e["reinit"].items()}, e.get("lr_change"))
PY
