cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== numerics"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x 2>&1 | tail -4
echo "== microbench"
timeout 600 python tools/microbench.py > gpurun_out/microbench_r2_call24.txt 2>&1; grep -i "8192 fwd" gpurun_out/microbench_r2_call24.txt | head -3
timeout 300 python tools/microbench_conv.py --out gpurun_out/microbench_conv_r2_call24.json 2>&1 | tail -1
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-400
echo "== bert"; timeout 600 python tools/bench_bert.py 2>&1 | tail -2 | cut -c1-400
