cd $GRAFT_REPO_ROOT
echo "== default"; timeout 300 python tools/dev/debug_write_sem.py 2>&1 | tail -26
echo "== BN_BULK=0"; MXB200_BN_BULK=0 timeout 300 python tools/dev/debug_write_sem.py 2>&1 | tail -12
echo "== gpu io"; timeout 600 python -m pytest tests/test_gpu_io.py -q 2>&1 | tail -5
echo "== bench_io"; timeout 600 python tools/bench_io.py --out gpurun_out/bench_io_r2_call17.json 2>&1 | tail -3
