cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for E in 0 1; do
  MXB200_BN_REDUCE_GRID=$E timeout 200 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -k regex:bwd_reduce_bulk -c 96 --csv --log-file gpurun_out/bn_reduce_grid${E}_r2_call33.csv python tools/microbench_bn.py > gpurun_out/bn33_$E.log 2>&1
  python - <<PY
import csv, re
from collections import defaultdict
lines=[l for l in open("gpurun_out/bn_reduce_grid${E}_r2_call33.csv") if not l.startswith("==")]
rows=defaultdict(dict)
for r in csv.DictReader(lines):
    rid=int(r["ID"]); v=float(r["Metric Value"].replace(",",""))
    if r["Metric Name"]=="gpu__time_duration.sum": rows[rid]["us"]=v*{"ns":1e-3,"nsecond":1e-3,"us":1,"usecond":1}.get(r.get("Metric Unit","ns"),1e-3)
    else: rows[rid]["grid"]=int(v)
ids=sorted(rows); out=[]
for k in range(0,len(ids),12):
    grp=[rows[i] for i in ids[k:k+12]]
    if grp: out.append("%d:%.1f" % (grp[-1].get("grid",0), sorted(g["us"] for g in grp)[len(grp)//2]))
print("MXB200_BN_REDUCE_GRID=${E} per shape (grid:median us):", " ".join(out))
PY
done
for E in 0 1; do echo "== bench GRID=$E"; MXB200_BN_REDUCE_GRID=$E timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-200; done
echo "== bn + resnet tests with GRID=1"; MXB200_BN_REDUCE_GRID=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "bn or resnet or gluon or module" 2>&1 | tail -3
