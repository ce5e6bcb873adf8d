cd $GRAFT_REPO_ROOT
O=gpurun_out/c15; mkdir -p $O
P=29950
for u in 0 1; do
P=$((P+1)); timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P tools/bandwidth/sweep.py --min-bytes 4194304 --no-graph --unroll $u --out $O/bandwidth_sweep_n8_unroll$u.json > $O/sweep_n8_u$u.txt 2>&1; echo "sweep8 unroll=$u rc=$?"
done
python - <<'PY'
import json
for u in (0, 1):
    try:
        d = json.load(open("gpurun_out/c15/bandwidth_sweep_n8_unroll%d.json" % u))
        for r in d["rows"]:
            print("unroll", u, r["bytes"], {k: round(v, 1) for k, v in r.items() if k.endswith("_us")}, {k: round(v, 2) for k, v in r.items() if k.endswith("roofline")}, [k for k, v in r.items() if k.endswith("_ok") and not v])
    except Exception as e:
        print("no sweep json", u, e)
PY
timeout 200 python tools/bandwidth/measure.py --gpus 0,1,2,3,4,5,6,7 --network resnet --num-layers 200 --num-batches 6 --mode kv --json $O/bandwidth_kv_singleproc_n8.json 2>&1 | tail -5
