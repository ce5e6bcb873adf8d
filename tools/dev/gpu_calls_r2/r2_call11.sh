cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/c11
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "bn or resnet or gluon or module" 2>&1 | tail -6
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c11/bench_graph.json 2> gpurun_out/c11/bench.err; cat gpurun_out/c11/bench_graph.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])"; tail -n 3 gpurun_out/c11/bench.err
timeout 900 python -m pytest tests/test_multigpu.py -q > gpurun_out/c11/pytest_multigpu.txt 2>&1; echo "pytest multigpu rc=$?"
tail -25 gpurun_out/c11/pytest_multigpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 tools/bandwidth/sweep.py --max-bytes 4194304 --out gpurun_out/c11/bandwidth_sweep_n2.json > gpurun_out/c11/sweep_n2.txt 2>&1; echo "sweep rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c11/bandwidth_sweep_n2.json"))
    for r in d["rows"]:
        print(r["bytes"], {k: round(v, 1) for k, v in r.items() if k.endswith("_us")}, {k: v for k, v in r.items() if k.endswith("_ok") and not v})
except Exception as e:
    print("no sweep json", e)
PY
tail -5 gpurun_out/c11/sweep_n2.txt
timeout 300 python tools/bandwidth/measure.py --gpus 0,1 --network resnet --num-layers 200 --num-batches 6 --mode kv --json gpurun_out/c11/bandwidth_kv_singleproc_n2.json 2>&1 | tail -7
mkdir -p gpurun_out/c11/elastic && cd gpurun_out/c11/elastic && rm -f host_worker* && \
  timeout 900 python $GRAFT_REPO_ROOT/tools/launch.py -n 1 -s 1 --launcher local --elastic-training-enabled True \
    python $GRAFT_REPO_ROOT/tools/elastic_gluon_worker.py $GRAFT_REPO_ROOT/gpurun_out/c11/elastic --plan 1,2,1 --layers 50 --image 224 --classes 1000 --batch-per-gpu 256 --steps 10 --lr 0.02 > launch.log 2>&1; echo "elastic rc=$?"
cd $GRAFT_REPO_ROOT; tail -5 gpurun_out/c11/elastic/launch.log; python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c11/elastic/out_*.json")):
    d = json.load(open(f))
    for e in d["epochs"]:
        print(d["tag"], {k: (round(v, 3) if isinstance(v, float) else v) for k, v in e.items() if k not in ("reinit", "lr_change")})
        if "reinit" in e: print("   reinit", {k: round(v, 1) if isinstance(v, float) else v for k, v in e["reinit"].items()}, e.get("lr_change"))
PY
