cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/c13
timeout 900 python -m pytest tests/test_multigpu.py -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 tools/bandwidth/sweep.py --min-bytes 1048576 --max-bytes 1073741824 --no-graph --out gpurun_out/c13/bandwidth_sweep_n2.json > gpurun_out/c13/sweep_n2.txt 2>&1; echo "sweep rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/c13/bandwidth_sweep_n2.json"))
for r in d["rows"]:
    print(r["bytes"], {k: round(v, 1) for k, v in r.items() if k.endswith("_us")}, {k: round(v, 2) for k, v in r.items() if k.endswith("roofline")}, [k for k, v in r.items() if k.endswith("_ok") and not v])
PY
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "write_gradient_semantics or fused_attention or fused_allreduce" 2>&1 | tail -4
