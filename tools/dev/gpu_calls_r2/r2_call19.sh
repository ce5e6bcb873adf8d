cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== write semantics debug"; timeout 300 python tools/dev/debug_write_sem.py 2>&1 | tail -26
echo "== conv shift numerics (mode 1, 2)"
MXB200_CONV_SHIFT=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "conv" 2>&1 | tail -3
MXB200_CONV_SHIFT=2 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "conv" 2>&1 | tail -3
echo "== conv microbench shift=1"; MXB200_CONV_SHIFT=1 timeout 300 python tools/microbench_conv.py --out gpurun_out/microbench_conv_shift1_r2_call19.json 2>&1 | tail -1
echo "== conv microbench shift=2"; MXB200_CONV_SHIFT=2 timeout 300 python tools/microbench_conv.py --out gpurun_out/microbench_conv_shift2_r2_call19.json 2>&1 | tail -1
echo "== bench_io backends"
timeout 300 python tools/bench_io.py --threads 32 --batches 16 2>&1 | tail -1
MXB200_NVJPEG_BACKEND=2 timeout 300 python tools/bench_io.py --threads 32 --batches 16 2>&1 | tail -1
echo "== ncu"
bash tools/dev/r2_call18.sh 2>&1 | tail -12
