cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/c6
P=29800
for net in dense resnet; do for nv in 1 0; do for sd in 1 0; do
P=$((P+1))
MXB200_NVLS=$nv MXB200_SEQ_DEV=$sd timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P tools/dev/diverge.py $net 2>&1 | grep -E "RESULT|Error|error" | head -5
done; done; done
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "conv3x3 or conv_generic or conv_nhwc_autograd" 2>&1 | tail -15
