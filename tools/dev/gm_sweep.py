import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
code = r'''
import torch, sys, os
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
out = {}
for (M, N, Kd) in [(8192, 8192, 8192), (16384, 8192, 4096), (8192, 3072, 768), (8192, 768, 3072), (50176, 512, 1024)]:
    x = torch.randn(M, Kd, device=dev).bfloat16(); w = torch.randn(N, Kd, device=dev).bfloat16()
    ours = t(lambda: K.gemm_nt(x, w)); ref = t(lambda: torch.matmul(x, w.t()))
    out["%dx%dx%d" % (M, N, Kd)] = (round(ours * 1e3, 1), round(ref * 1e3, 1), round(ref / ours, 3))
print(out)
'''
for gm in ["", "2", "4", "6", "8", "12", "32"]:
    env = dict(os.environ)
    if gm:
        env["MXB200_GEMM_GM"] = gm
    r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print("gm=%s" % (gm or "default16"), r.stdout.strip().splitlines()[-1][:400], flush=True)
