set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/c3
timeout 600 python tools/microbench_conv.py --out gpurun_out/c3/microbench_conv.json > gpurun_out/c3/microbench_conv.txt 2>&1; echo "rc=$?"
cat gpurun_out/c3/microbench_conv.txt | python -c "
import sys, json
for ln in sys.stdin:
    try: r = json.loads(ln)
    except Exception: print(ln.strip()); continue
    if 'shape' in r:
        print('%-28s fwd %7.1f us (%.2fx cudnn, %4.0f TF) dgrad %7.1f (%.2fx, %4.0f TF) wgrad %7.1f (%.2fx, %4.0f TF)' % (r['shape'], r['fwd_us'], r['fwd_vs_cudnn'], r['fwd_tflops'], r['dgrad_us'], r['dgrad_vs_cudnn'], r['dgrad_tflops'], r['wgrad_us'], r['wgrad_vs_cudnn'], r['wgrad_tflops']))
    else: print(r)
"
timeout 600 python tools/dev/check_graph.py > gpurun_out/c3/check_graph.txt 2>&1; echo "check_graph rc=$?"
tail -3 gpurun_out/c3/check_graph.txt
