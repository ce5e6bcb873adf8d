#!/usr/bin/env python
"""Single-kernel driver for ncu: runs one GEMM / BN shape a few times (ncu -k regex:... -c N selects launches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxb200.ops import functional as K  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "gemm_fwd"
dev = torch.device("cuda", 0)
if what.startswith("gemm"):
    M, N, Kd = [int(v) for v in (sys.argv[2:5] if len(sys.argv) >= 5 else (802816, 256, 64))]
    x = torch.randn(M, Kd, device=dev).bfloat16()
    w = torch.randn(N, Kd, device=dev).bfloat16()
    dy = torch.randn(M, N, device=dev).bfloat16()
    for _ in range(4):
        if what == "gemm_fwd":
            K.gemm_nt(x, w)
        elif what == "gemm_dgrad":
            K.gemm_nn(dy, w)
        else:
            K.gemm_tn(dy, x, out_dtype=torch.float32)
else:
    M, C = [int(v) for v in sys.argv[2:4]] if len(sys.argv) >= 4 else (802816, 256)
    x = torch.randn(M, C, device=dev).bfloat16().requires_grad_(True)
    r = torch.randn(M, C, device=dev).bfloat16().requires_grad_(True)
    g, b = torch.ones(C, device=dev, requires_grad=True), torch.zeros(C, device=dev, requires_grad=True)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    for _ in range(3):
        y = K.bn_act(x, g, b, rm, rv, residual=r, relu=True)
        y.backward(torch.ones_like(y))
torch.cuda.synchronize()
