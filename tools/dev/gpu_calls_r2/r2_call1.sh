set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/c1
timeout 600 python tools/dev/check_graph.py > gpurun_out/c1/check_graph.txt 2>&1; echo "check_graph rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c1/bench_graph.json 2> gpurun_out/c1/bench_graph.err; echo "rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-graph > gpurun_out/c1/bench_eager.json 2> gpurun_out/c1/bench_eager.err; echo "rc=$?"
timeout 600 python bench.py --impl library --steps 20 --warmup 5 > gpurun_out/c1/bench_library.json 2> gpurun_out/c1/bench_library.err; echo "rc=$?"
tail -3 gpurun_out/c1/check_graph.txt; cat gpurun_out/c1/bench_*.json; tail -5 gpurun_out/c1/*.err
