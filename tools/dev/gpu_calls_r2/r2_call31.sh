cd $GRAFT_REPO_ROOT
echo "== numerics with pair128"; MXB200_GEMM_2CTA_128=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "gemm or linear or conv1x1 or resnet" 2>&1 | tail -3
for E in 0 1; do
echo "== MXB200_GEMM_2CTA_128=$E"
MXB200_GEMM_2CTA_128=$E timeout 300 python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.environ["GRAFT_REPO_ROOT"])
from mxb200.ops import functional as K
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, it=12):
    for _ in range(4): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(it):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
for (M, N, Kd) in [(802816, 128, 256), (200704, 128, 512), (8192, 768, 3072), (8192, 768, 768), (12544, 512, 2048), (8192, 8192, 8192)]:
    x = torch.randn(M, Kd, device=dev).bfloat16(); w = torch.randn(N, Kd, device=dev).bfloat16(); dy = torch.randn(M, N, device=dev).bfloat16()
    a = t(lambda: K.gemm_nt(x, w, block_n=128)); d = t(lambda: K.gemm_nn(dy, w, block_n=128)); r = t(lambda: torch.matmul(x, w.t()))
    y1 = K.gemm_nt(x, w, block_n=128); err = ((y1.float() - x.float() @ w.float().t()).abs().max() / (x.float() @ w.float().t()).abs().max()).item() if M * N < 5e7 else -1
    print("%6d x %5d x %5d  fwd(bn128) %.1f us  dgrad(bn128) %.1f us   cublas fwd %.1f   relerr %.3g" % (M, N, Kd, a * 1e3, d * 1e3, r * 1e3, err))
PY
done
