#!/usr/bin/env python
"""Numerics + timing of the CTA-pair (cta_group::2) GEMM path.  Run with MXB200_GEMM_2CTA=1 (forced) or =0 (off);
prints one JSON line: max relative errors per case and ms per shape (CUDA events, L2 flushed)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxb200.ops import functional as K  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-6)).item()


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


out = {"mode": os.environ.get("MXB200_GEMM_2CTA", "auto"), "errors": {}, "ms": {}}
torch.manual_seed(0)
for (M, N, Kd) in [(512, 256, 64), (2500, 520, 200), (4224, 1024, 256), (20000, 256, 512), (300, 256, 128)]:
    x = torch.randn(M, Kd, device=dev).bfloat16()
    w = (torch.randn(N, Kd, device=dev) * 0.1).bfloat16()
    b = torch.randn(N, device=dev)
    y = K.gemm_nt(x, w, bias=b, relu=True, block_n=256)
    e1 = rel(y, torch.relu(x.float() @ w.float().t() + b))
    s, s2 = torch.zeros(N, device=dev), torch.zeros(N, device=dev)
    y = K.gemm_nt(x, w, col_stats=(s, s2), block_n=256)
    e2 = max(rel(y, x.float() @ w.float().t()), rel(s, y.float().sum(0)), rel(s2, (y.float() ** 2).sum(0)))
    dy = (torch.randn(M, N, device=dev) * 0.1).bfloat16()
    add = torch.randn(M, Kd, device=dev).bfloat16()
    dx = K.gemm_nn(dy, w, addend=add, block_n=256 if Kd >= 256 else 0)
    e3 = rel(dx, dy.float() @ w.float() + add.float())
    out["errors"]["%dx%dx%d" % (M, N, Kd)] = [e1, e2, e3]
for (M, N, Kd) in [(8192, 8192, 8192), (50176, 1024, 256), (50176, 256, 1024), (50176, 512, 1024), (12544, 2048, 512),
                   (12544, 512, 2048), (8192, 2304, 768), (8192, 3072, 768), (8192, 768, 3072), (200704, 256, 512)]:
    x = torch.randn(M, Kd, device=dev).bfloat16()
    w = torch.randn(N, Kd, device=dev).bfloat16()
    dy = torch.randn(M, N, device=dev).bfloat16()
    t_f = timeit(lambda: K.gemm_nt(x, w))
    t_d = timeit(lambda: K.gemm_nn(dy, w))
    t_c = timeit(lambda: torch.matmul(x, w.t()))
    t_w = timeit(lambda: K.gemm_tn(dy, x, out_dtype=torch.float32)) if M <= 250000 else float("nan")
    if M <= 60000:
        e_w = rel(K.gemm_tn(dy, x, out_dtype=torch.float32), dy.float().t() @ x.float())
        out["errors"]["wgrad_%dx%dx%d" % (M, N, Kd)] = [e_w]
    fl = 2.0 * M * N * Kd
    out["ms"]["%dx%dx%d" % (M, N, Kd)] = {"fwd": t_f, "dgrad": t_d, "cublas_fwd": t_c, "wgrad": t_w, "fwd_tflops": fl / t_f / 1e9, "cublas_tflops": fl / t_c / 1e9}
    del x, w, dy
print(json.dumps(out))
