#!/usr/bin/env python
"""Kernel micro-benchmarks on one B200: our tcgen05 GEMM vs cuBLAS (torch.matmul), fused BN(+add)+ReLU vs
torch eager, and roofline fractions against MEASURED_PEAKS.json.  CUDA-event timed, warm-up 5, L2 flushed
between iterations by writing a 256 MB buffer."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxb200.ops import functional as K  # noqa: E402

dev = torch.device("cuda", 0)
PEAKS = {"hbm_gbs": 6583.8, "bf16_tflops": 1710.4}
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
        PEAKS.update(json.load(f))
except Exception:  # noqa: BLE001
    pass
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


results = []
print("== GEMM (ResNet-50 batch 256 1x1 convs + FC), ms / TFLOP/s / frac of measured cuBLAS peak")
B = 256
shapes = [  # (M pixels, Cout, Cin)
    (B * 56 * 56, 64, 64), (B * 56 * 56, 256, 64), (B * 56 * 56, 64, 256), (B * 56 * 56, 128, 256),
    (B * 28 * 28, 512, 128), (B * 28 * 28, 128, 512), (B * 28 * 28, 256, 512),
    (B * 14 * 14, 1024, 256), (B * 14 * 14, 256, 1024), (B * 14 * 14, 512, 1024),
    (B * 7 * 7, 2048, 512), (B * 7 * 7, 512, 2048), (B, 1000, 2048), (8192, 8192, 8192),
]
for (M, N, Kd) in shapes:
    x = torch.randn(M, Kd, device=dev).bfloat16()
    w = torch.randn(N, Kd, device=dev).bfloat16()
    dy = torch.randn(M, N, device=dev).bfloat16()
    fl = 2.0 * M * N * Kd
    row = {"M": M, "N": N, "K": Kd}
    for name, ours, ref in [
        ("fwd", lambda: K.gemm_nt(x, w), lambda: torch.matmul(x, w.t())),
        ("dgrad", lambda: K.gemm_nn(dy, w), lambda: torch.matmul(dy, w)),
        ("wgrad", lambda: K.gemm_tn(dy, x, out_dtype=torch.float32), lambda: torch.matmul(dy.t(), x)),
    ]:
        if M == 8192 and name != "fwd":
            continue
        t_o, t_r = timeit(ours), timeit(ref)
        row[name] = {"ours_ms": t_o, "cublas_ms": t_r, "ours_tflops": fl / t_o / 1e9, "cublas_tflops": fl / t_r / 1e9,
                     "frac_of_peak": fl / t_o / 1e9 / PEAKS["bf16_tflops"]}
        print("M=%7d N=%5d K=%5d %-5s ours %.3f ms (%.0f TF, %.2f of peak)  cublas %.3f ms (%.0f TF)  ratio %.2f" % (
            M, N, Kd, name, t_o, fl / t_o / 1e9, fl / t_o / 1e9 / PEAKS["bf16_tflops"], t_r, fl / t_r / 1e9, t_r / t_o))
    results.append(row)
    del x, w, dy

print("== GEMM tile-width sweep (block_n 128 vs 256, fwd with and without fused BN statistics)")
sweep = []
for (M, N, Kd) in [s for s in shapes if s[1] >= 256 and s[0] > 256]:
    x = torch.randn(M, Kd, device=dev).bfloat16()
    w = torch.randn(N, Kd, device=dev).bfloat16()
    dy = torch.randn(M, N, device=dev).bfloat16()
    st = (torch.zeros(N, device=dev), torch.zeros(N, device=dev))
    row = {"M": M, "N": N, "K": Kd}
    for bn in (128, 256):
        row["fwd_bn%d" % bn] = timeit(lambda: K.gemm_nt(x, w, block_n=bn))
        row["fwd_stats_bn%d" % bn] = timeit(lambda: K.gemm_nt(x, w, col_stats=st, block_n=bn))
        if M != 8192:
            row["dgrad_bn%d" % bn] = timeit(lambda: K.gemm_nn(dy, w, block_n=bn if Kd >= 256 else 0))
            if Kd >= 256:
                row["wgrad_bn%d" % bn] = timeit(lambda: K.gemm_tn(dy, x, out_dtype=torch.float32, block_n=bn))
    row["fwd_auto"] = timeit(lambda: K.gemm_nt(x, w))
    print(" ".join("%s=%s" % (k, ("%.3f" % v) if isinstance(v, float) else v) for k, v in row.items()))
    sweep.append(row)
    del x, w, dy

print("== BN(+add)+ReLU training fwd+bwd, NHWC bf16: ours vs torch eager (batch_norm + add + relu, channels_last)")
bn_rows = []
for (n, h, c, res) in [(256, 112, 64, False), (256, 56, 64, False), (256, 56, 256, True), (256, 28, 512, True),
                       (256, 14, 1024, True), (256, 7, 2048, True)]:
    M = n * h * h
    x = torch.randn(M, c, device=dev).bfloat16().requires_grad_(True)
    r = torch.randn(M, c, device=dev).bfloat16().requires_grad_(True) if res else None
    g, b = torch.ones(c, device=dev, requires_grad=True), torch.zeros(c, device=dev, requires_grad=True)
    rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    gy = torch.randn(M, c, device=dev).bfloat16()

    def ours():
        y = K.bn_act(x, g, b, rm, rv, residual=r, relu=True)
        y.backward(gy)
        x.grad = None

    x4 = x.detach().reshape(n, h, h, c).permute(0, 3, 1, 2).requires_grad_(True)
    r4 = r.detach().reshape(n, h, h, c).permute(0, 3, 1, 2).requires_grad_(True) if res else None
    gy4 = gy.reshape(n, h, h, c).permute(0, 3, 1, 2)
    gb, bb = g.detach().bfloat16().requires_grad_(True), b.detach().bfloat16().requires_grad_(True)

    def ref():
        y = torch.nn.functional.batch_norm(x4, rm, rv, gb, bb, True, 0.1, 1e-5)
        if res:
            y = y + r4
        y = torch.relu(y)
        y.backward(gy4)
        x4.grad = None

    t_o, t_r = timeit(ours), timeit(ref)
    # algorithmic bytes: fwd read x(2x: stats+apply)+res, write y ; bwd read dy,x,y twice, write dx(+dres)
    elems = M * c
    byts = 2 * elems * ((2 + (1 if res else 0) + 1) + (6 + 1 + (1 if res else 0)))
    bn_rows.append({"M": M, "C": c, "res": res, "ours_ms": t_o, "torch_ms": t_r, "ours_gbs": byts / t_o / 1e6,
                    "frac_of_hbm": byts / t_o / 1e6 / PEAKS["hbm_gbs"]})
    print("M=%8d C=%5d res=%d  ours %.3f ms (%.0f GB/s, %.2f of measured HBM)  torch %.3f ms  speedup %.2f" % (
        M, c, res, t_o, byts / t_o / 1e6, byts / t_o / 1e6 / PEAKS["hbm_gbs"], t_r, t_r / t_o))
    del x, r, x4, r4, gy

os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/microbench.json", "w") as f:
    json.dump({"peaks": PEAKS, "gemm": results, "gemm_block_n_sweep": sweep, "bn": bn_rows}, f, indent=1)
