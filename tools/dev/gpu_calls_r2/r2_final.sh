cd $GRAFT_REPO_ROOT
O=gpurun_out/final; mkdir -p $O
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 | tee $O/bench_reference_1gpu.json | cut -c1-200
echo "== full gpu suite"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench ours"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee $O/bench_ours_1gpu.json | cut -c1-260
echo "== bench library"; timeout 600 python bench.py --impl library --steps 10 --warmup 3 2>&1 | tail -1 | tee $O/bench_library_1gpu.json | cut -c1-260
