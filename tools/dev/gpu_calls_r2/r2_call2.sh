set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/c2
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "conv3x3 or conv_generic or conv_nhwc_autograd" > gpurun_out/c2/pytest_conv.txt 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/c2/pytest_conv.txt
timeout 600 python tools/dev/check_graph.py > gpurun_out/c2/check_graph.txt 2>&1; echo "check_graph rc=$?"
tail -3 gpurun_out/c2/check_graph.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c2/bench_graph.json 2> gpurun_out/c2/bench_graph.err; echo "rc=$?"
cat gpurun_out/c2/bench_graph.json; tail -n 5 gpurun_out/c2/bench_graph.err
