cd $GRAFT_REPO_ROOT
echo "== full gpu suite"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench ours"; timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_confirm.json | cut -c1-200
