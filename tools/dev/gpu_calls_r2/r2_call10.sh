cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/c10
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "bn or conv3x3_row_shifted or resnet" 2>&1 | tail -6
echo "== bulk reduce"; timeout 300 python tools/microbench_bn.py gpurun_out/c10/microbench_bn_bulk.json | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print('%8d x %4d  fwd %7.1f us (%.2f of hbm)  bwd %7.1f us (%.2f of hbm)' % (r['M'], r['C'], r['fwd_stats_apply_us'], r['fwd_frac_of_measured_hbm'], r['bwd_reduce_apply_us'], r['bwd_frac_of_measured_hbm']))"
echo "== register-prefetch reduce"; MXB200_BN_BULK=0 timeout 300 python tools/microbench_bn.py gpurun_out/c10/microbench_bn_old.json | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print('%8d x %4d  fwd %7.1f us (%.2f of hbm)  bwd %7.1f us (%.2f of hbm)' % (r['M'], r['C'], r['fwd_stats_apply_us'], r['fwd_frac_of_measured_hbm'], r['bwd_reduce_apply_us'], r['bwd_frac_of_measured_hbm']))"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c10/bench_graph.json 2> gpurun_out/c10/bench.err; cat gpurun_out/c10/bench_graph.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])"; tail -n 3 gpurun_out/c10/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bwd_reduce_bulk -s 20 -c 2 -o gpurun_out/c10/ncu_bn_bwd_reduce_bulk python tools/microbench_bn.py > /dev/null 2>&1; echo "ncu rc=$?"
