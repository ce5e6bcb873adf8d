cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/c7
for bo in 1 0; do
echo "=== MXB200_SHIFT_BASEOFF=$bo"
MXB200_SHIFT_BASEOFF=$bo timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "conv3x3 or conv_generic or conv_nhwc_autograd" 2>&1 | tail -8
done
echo "=== attention"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "fused_attention" 2>&1 | tail -25
