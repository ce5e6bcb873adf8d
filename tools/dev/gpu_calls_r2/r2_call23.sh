cd $GRAFT_REPO_ROOT
echo "== reference arm (3 steps)"; timeout 1200 python bench.py --impl reference --steps 3 --warmup 3 2>&1 | tail -3
echo "== reference direct stderr tail"; LD_LIBRARY_PATH=baseline/_ref/mxnet timeout 600 python baseline/run_reference.py --gpus 1 --steps 2 --warmup 3 --batch 64 2>&1 | tail -6
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
