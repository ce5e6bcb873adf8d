cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv --log-file gpurun_out/launches_r2_call32.csv python bench.py --no-graph --steps 2 --warmup 3 > gpurun_out/launches_r2_call32.log 2>&1
tail -2 gpurun_out/launches_r2_call32.log | cut -c1-200; wc -l gpurun_out/launches_r2_call32.csv
python tools/launch_table.py gpurun_out/launches_r2_call32.csv | head -50
