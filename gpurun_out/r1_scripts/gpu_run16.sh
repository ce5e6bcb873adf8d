#!/bin/bash
# 1-GPU: CTA-pair GEMM check (forced on / off / auto), then tests + bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
MXB200_GEMM_2CTA=1 timeout 300 python tools/check_gemm_2cta.py > gpurun_out/gemm2cta_on.json 2> gpurun_out/gemm2cta_on.err; echo "2cta on exit $?"; tail -3 gpurun_out/gemm2cta_on.err | cut -c1-300
MXB200_GEMM_2CTA=0 timeout 300 python tools/check_gemm_2cta.py > gpurun_out/gemm2cta_off.json 2> gpurun_out/gemm2cta_off.err; echo "2cta off exit $?"
python - <<'PY'
import json
try:
    on=json.load(open("gpurun_out/gemm2cta_on.json")); off=json.load(open("gpurun_out/gemm2cta_off.json"))
    print("errors(2cta):", {k:[round(e,4) for e in v] for k,v in on["errors"].items()})
    for k in on["ms"]:
        a,b=on["ms"][k],off["ms"][k]
        print("%-18s fwd 2cta %.3f ms (%.0f TF) | 1cta %.3f ms (%.0f TF) | cublas %.3f (%.0f TF) | dgrad 2cta %.3f 1cta %.3f" % (k,a["fwd"],a["fwd_tflops"],b["fwd"],b["fwd_tflops"],b["cublas_fwd"],b["cublas_tflops"],a["dgrad"],b["dgrad"]))
except Exception as e:
    print("compare failed", e)
PY
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t16_all.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/t16_all.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench16.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench16.log | cut -c1-300
timeout 300 python tools/bench_bert.py --steps 10 --warmup 3 > gpurun_out/bert16.log 2>&1; echo "bert exit $?"; tail -1 gpurun_out/bert16.log | cut -c1-300
timeout 600 python examples/image-classification/benchmark_score.py --networks resnet50_v1b --batch-sizes 32,128 --dtype bfloat16 > gpurun_out/score16.log 2>&1; grep "images/sec" gpurun_out/score16.log
