cd $GRAFT_REPO_ROOT
O=gpurun_out/c12; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
P=29900
run_bench() {  # n impl extra
  P=$((P+1))
  if [ "$1" = "1" ]; then timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --impl $2 $3 > $O/bench_$2_n$1.json 2> $O/bench_$2_n$1.err
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $P bench.py --gpus $1 --steps 20 --warmup 5 --impl $2 $3 > $O/bench_$2_n$1.json 2> $O/bench_$2_n$1.err; fi
  echo "bench $2 n=$1 rc=$?"; python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/bench_$2_n$1.json") if l.startswith("{")][-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d.get("e2e", {}).get("value"), d.get("step_ms", {}).get("per_rank_ms_min_med_max"))
except Exception as e:
    print("no json", e)
PY
}
run_bench 8 ours
run_bench 4 ours
run_bench 2 ours
run_bench 1 ours
run_bench 8 ours
run_bench 8 library
P=$((P+1)); timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P tools/bandwidth/sweep.py --out $O/bandwidth_sweep_n8.json > $O/sweep_n8.txt 2>&1; echo "sweep8 rc=$?"
P=$((P+1)); timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $P tools/bandwidth/sweep.py --max-bytes 268435456 --out $O/bandwidth_sweep_n4.json > $O/sweep_n4.txt 2>&1; echo "sweep4 rc=$?"
python - <<'PY'
import json
for n in (8, 4):
    try:
        d = json.load(open("gpurun_out/c12/bandwidth_sweep_n%d.json" % n))
        for r in d["rows"]:
            print(n, r["bytes"], {k: round(v, 1) for k, v in r.items() if k.endswith("_us")}, {k: round(v, 2) for k, v in r.items() if k.endswith("roofline")}, [k for k, v in r.items() if k.endswith("_ok") and not v])
    except Exception as e:
        print("no sweep json", n, e)
PY
timeout 600 python -m pytest tests/test_multigpu.py -q > $O/pytest_multigpu_n8.txt 2>&1; echo "pytest multigpu rc=$?"; tail -5 $O/pytest_multigpu_n8.txt
timeout 300 python tools/bandwidth/measure.py --gpus 0,1,2,3,4,5,6,7 --network resnet --num-layers 200 --num-batches 6 --mode kv --json $O/bandwidth_kv_singleproc_n8.json 2>&1 | tail -6
P=$((P+1)); timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P tools/bench_bert.py --steps 20 --warmup 5 > $O/bert_8gpu.json 2> $O/bert_8gpu.err; echo "bert8 rc=$?"; tail -n 1 $O/bert_8gpu.json
mkdir -p $O/elastic && cd $O/elastic && rm -f host_worker* && \
  timeout 900 python $GRAFT_REPO_ROOT/tools/launch.py -n 4 -s 1 --launcher local --elastic-training-enabled True \
    python $GRAFT_REPO_ROOT/tools/elastic_gluon_worker.py $GRAFT_REPO_ROOT/$O/elastic --plan 4,8,4 --layers 50 --image 224 --classes 1000 --batch-per-gpu 512 --steps 12 --lr 0.02 > launch.log 2>&1; echo "elastic rc=$?"
cd $GRAFT_REPO_ROOT; tail -4 $O/elastic/launch.log; python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c12/elastic/out_*.json")):
    d = json.load(open(f))
    for e in d["epochs"]:
        if d["tag"] in ("local0", "local4"):
            print(d["tag"], {k: (round(v, 3) if isinstance(v, float) else v) for k, v in e.items() if k not in ("reinit", "lr_change")})
            if "reinit" in e: print("   reinit", {k: round(v, 1) if isinstance(v, float) else v for k, v in e["reinit"].items()}, e.get("lr_change"))
PY
