cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/c9
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c9/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/c9/pytest_gpu.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 2000 --csv --log-file gpurun_out/c9/launches_bert.csv python tools/bench_bert.py --steps 4 --warmup 4 > gpurun_out/c9/ncu_bert.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/c9/launches_bert.csv > gpurun_out/c9/launch_breakdown_bert.txt 2>&1; head -32 gpurun_out/c9/launch_breakdown_bert.txt
python -c "
import __graft_entry__ as g
g.smoke()" 2>&1 | tail -3
