cd $GRAFT_REPO_ROOT
O=gpurun_out/c26; mkdir -p $O
echo "== fold test in sequence + storage + gpu io"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_storage.py tests/test_gpu_io.py -q -m gpu -x 2>&1 | tail -4
echo "== multigpu tests"; timeout 900 python -m pytest tests/test_multigpu.py -q > $O/pytest_multigpu.txt 2>&1; echo "rc=$?"; tail -6 $O/pytest_multigpu.txt
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | tail -1 | cut -c1-330
echo "== kv per-key N=2"; timeout 300 python tools/bandwidth/measure.py --gpus 0,1 --network resnet --num-layers 200 --num-batches 6 --mode kv --json $O/bandwidth_kv_singleproc_n2.json 2>&1 | tail -4
echo "== elastic 1-2-1"
mkdir -p $O/elastic && cd $O/elastic && rm -f host_worker* && \
  timeout 900 python $GRAFT_REPO_ROOT/tools/launch.py -n 1 -s 1 --launcher local --elastic-training-enabled True \
    python $GRAFT_REPO_ROOT/tools/elastic_gluon_worker.py $GRAFT_REPO_ROOT/$O/elastic --plan 1,2,1 --layers 50 --image 224 --classes 1000 --batch-per-gpu 256 --steps 10 --lr 0.02 > launch.log 2>&1; echo "elastic rc=$?"
cd $GRAFT_REPO_ROOT; tail -3 $O/elastic/launch.log; python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c26/elastic/out_*.json")):
    d = json.load(open(f))
    for e in d["epochs"]:
        print(d["tag"], {k: (round(v, 3) if isinstance(v, float) else v) for k, v in e.items() if k not in ("reinit", "lr_change")})
        if "reinit" in e: print("   reinit", {k: round(v, 1) if isinstance(v, float) else v for k, v in e["reinit"].items()}, e.get("lr_change"))
PY
