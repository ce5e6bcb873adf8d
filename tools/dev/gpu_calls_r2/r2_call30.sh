cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm" 2>&1 | tail -2
timeout 600 python tools/microbench.py > gpurun_out/microbench_r2_call30.txt 2>&1; grep -E "K= (2048|1024|512) (fwd|dgrad)|8192 fwd" gpurun_out/microbench_r2_call30.txt | head -20
MXB200_GEMM_GM=16 timeout 300 python - <<'PY'
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
for (M, N, Kd) in [(8192, 768, 3072), (8192, 768, 768), (8192, 3072, 768), (8192, 2304, 768), (12544, 512, 2048), (50176, 512, 1024)]:
    x = torch.randn(M, Kd, device=dev).bfloat16(); w = torch.randn(N, Kd, device=dev).bfloat16()
    a = t(lambda: K.gemm_nt(x, w)); b = t(lambda: K.gemm_nt(x, w, block_n=256)); c = t(lambda: K.gemm_nt(x, w, block_n=128)); r = t(lambda: torch.matmul(x, w.t()))
    print("%6d x %5d x %5d  auto %.1f us  bn256 %.1f  bn128 %.1f  cublas %.1f" % (M, N, Kd, a * 1e3, b * 1e3, c * 1e3, r * 1e3))
PY
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-260
echo "== bert"; timeout 600 python tools/bench_bert.py 2>&1 | tail -1 | cut -c1-200
