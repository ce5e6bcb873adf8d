cd $GRAFT_REPO_ROOT
echo "== failing test x2"
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "inference_folded" --tb=short 2>&1 | tail -25
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "inference_folded" --tb=line 2>&1 | tail -5
echo "== bert new attention"; timeout 600 python tools/bench_bert.py 2>&1 | tail -1 | cut -c1-200
echo "== bert old attention"; MXB200_CUDA_LIB=$GRAFT_REPO_ROOT/mxb200/libmxb200_cuda_alt.so timeout 600 python tools/bench_bert.py 2>&1 | tail -1 | cut -c1-200
echo "== bert new attention again"; timeout 600 python tools/bench_bert.py 2>&1 | tail -1 | cut -c1-200
