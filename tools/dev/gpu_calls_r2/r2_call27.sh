cd $GRAFT_REPO_ROOT
O=gpurun_out/c27; mkdir -p $O
echo "== bench N=8 ours"; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --steps 20 --warmup 5 2>$O/bench8.err | tail -1 > $O/bench_8gpu.json; cut -c1-330 $O/bench_8gpu.json; python -c "
import json; d=json.load(open('$O/bench_8gpu.json')); print('e2e', d['e2e']['value'], 'clocks', d['clocks'], 'step_ms', d['step_ms']['per_rank_ms_min_med_max'][:2])"
echo "== bench N=8 reference (2 steps)"; timeout 200 python bench.py --impl reference --gpus 8 --steps 2 --warmup 3 2>&1 | tail -1 | cut -c1-300
