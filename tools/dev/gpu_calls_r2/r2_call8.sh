cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/c8
timeout 600 python tools/microbench_conv.py --out gpurun_out/c8/microbench_conv.json > gpurun_out/c8/microbench_conv.txt 2>&1; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/c8/microbench_conv.json'))
for r in d['rows']:
    print('%-28s x%d fwd %7.1f us (%.2fx, cudnn %6.1f) dgrad %7.1f (%.2fx, %6.1f) wgrad %7.1f (%.2fx, %6.1f)' % (r['shape'], r['count'], r['fwd_us'], r['fwd_vs_cudnn'], r['cudnn_fwd_us'], r['dgrad_us'], r['dgrad_vs_cudnn'], r['cudnn_dgrad_us'], r['wgrad_us'], r['wgrad_vs_cudnn'], r['cudnn_wgrad_us']))
print(d['summary'])
PY
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c8/bench_graph.json 2> gpurun_out/c8/bench_graph.err; echo "rc=$?"
cat gpurun_out/c8/bench_graph.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e']['value'])"
tail -n 3 gpurun_out/c8/bench_graph.err
timeout 600 python tools/bench_bert.py --steps 20 --warmup 5 > gpurun_out/c8/bert_1gpu.json 2> gpurun_out/c8/bert.err; echo "rc=$?"; cat gpurun_out/c8/bert_1gpu.json; tail -n 3 gpurun_out/c8/bert.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 1300 --csv --log-file gpurun_out/c8/launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-e2e > gpurun_out/c8/ncu_bench.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/c8/launches.csv > gpurun_out/c8/launch_breakdown.txt 2>&1; head -40 gpurun_out/c8/launch_breakdown.txt
