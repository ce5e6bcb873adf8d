#!/usr/bin/env python
"""Single-kernel drivers for ncu (one GPU, a handful of launches; pick the launch with `ncu -k regex:... -s N -c 1`).

    python tools/prof_kernels.py cublas 8192 8192 8192        # torch.matmul (cuBLAS) -- the library yardstick
    python tools/prof_kernels.py conv_fwd|conv_dgrad|conv_wgrad N H W C O k stride
    python tools/prof_kernels.py attn B S heads               # fused attention fwd + bwd (head dim 64)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxb200.ops import functional as K  # noqa: E402

what = sys.argv[1]
a = [int(v) for v in sys.argv[2:]]
dev = torch.device("cuda", 0)
if what == "cublas":
    M, N, Kd = a or (8192, 8192, 8192)
    x = torch.randn(M, Kd, device=dev).bfloat16()
    w = torch.randn(N, Kd, device=dev).bfloat16()
    for _ in range(5):
        torch.matmul(x, w.t())
elif what.startswith("conv"):
    n, h, w_, c, o, k, s = a or (512, 14, 14, 256, 256, 3, 1)
    pad = k // 2
    x = torch.randn(n, h, w_, c, device=dev).bfloat16()
    wk = (torch.randn(o, k, k, c, device=dev) * 0.05).bfloat16()
    P = (h + 2 * pad - k) // s + 1
    dy = torch.randn(n, P, P, o, device=dev).bfloat16()
    for _ in range(5):
        if what == "conv_fwd":
            K.conv_fwd_raw(x, wk, s, pad)
        elif what == "conv_dgrad":
            K.conv_dgrad_raw(dy, wk, tuple(x.shape), s, pad)
        else:
            K.conv_wgrad_raw(dy, x, tuple(wk.shape), s, pad)
elif what == "attn":
    from mxb200.ops.attention import self_attention
    B, S, nh = a or (64, 128, 12)
    q, k, v = [torch.randn(B, S, nh * 64, device=dev).bfloat16().requires_grad_(True) for _ in range(3)]
    for _ in range(4):
        o = self_attention(q, k, v, nh, p_drop=0.1, training=True)
        o.backward(torch.ones_like(o))
torch.cuda.synchronize()
