#!/bin/bash
# Run the GPU test groups in separate processes (a device-side trap poisons only its own process).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MXB200_NO_BUILD=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import torch; print(torch.__version__, torch.cuda.get_device_name(0))" > gpurun_out/env.txt 2>&1
for grp in "gemm" "bn_act" "softmax" "sgd_kernel or adam or elementwise" "fused_allreduce" "linear_autograd"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "$grp" > "gpurun_out/test_${name}.log" 2>&1
  echo "group [$grp] exit $?" | tee -a gpurun_out/summary.txt
  tail -5 "gpurun_out/test_${name}.log"
done
