cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== targeted"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "write_gradient_semantics or gluon_fused or module_fused or fused_allreduce" 2>&1 | tail -4
echo "== gemm / conv numerics"
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "gemm or conv or resnet or attention" 2>&1 | tail -6
echo "== gpu io"
timeout 600 python -m pytest tests/test_gpu_io.py -q 2>&1 | tail -15
echo "== microbench"
timeout 600 python tools/microbench.py > gpurun_out/microbench_r2_call16.txt 2>&1; grep -i "8192\|^==" gpurun_out/microbench_r2_call16.txt | head; tail -3 gpurun_out/microbench_r2_call16.txt
timeout 300 python tools/microbench_conv.py --out gpurun_out/microbench_conv_r2_call16.json 2>&1 | tail -8
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -2
