#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "maxpool or gemm" > gpurun_out/t24.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/t24.log
MXB200_GEMM_2CTA_WGRAD=1 timeout 300 python tools/check_gemm_2cta.py > gpurun_out/gemm2cta_wgrad_on.json 2> gpurun_out/gemm2cta_wgrad_on.err; echo "wgrad 2cta on exit $?"; tail -2 gpurun_out/gemm2cta_wgrad_on.err | cut -c1-300
timeout 300 python tools/check_gemm_2cta.py > gpurun_out/gemm2cta_wgrad_off.json 2> gpurun_out/gemm2cta_wgrad_off.err; echo "wgrad 2cta off exit $?"
python - <<'PY'
import json
on=json.load(open("gpurun_out/gemm2cta_wgrad_on.json")); off=json.load(open("gpurun_out/gemm2cta_wgrad_off.json"))
print({k:[round(e,4) for e in v] for k,v in on["errors"].items() if k.startswith("wgrad")})
for k in on["ms"]:
    print("%-18s wgrad 2cta %.3f ms | 1cta %.3f ms" % (k, on["ms"][k]["wgrad"], off["ms"][k]["wgrad"]))
PY
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench24.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench24.log | cut -c1-300
