#!/usr/bin/env python
"""BatchNorm kernel micro-benchmark: forward apply, backward reduce + apply for the ResNet-50 layer shapes at the benchmark
batch, reported as achieved HBM bandwidth (algorithmic bytes / CUDA-event time) against MEASURED_PEAKS.json hbm_gbs."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from mxb200.ops import functional as K  # noqa: E402

dev = torch.device("cuda", 0)
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:  # noqa: BLE001
    pass
HBM = peaks.get("hbm_gbs", 6650.0)
SHAPES = [(512 * 56 * 56, 64), (512 * 56 * 56, 256), (512 * 28 * 28, 128), (512 * 28 * 28, 512), (512 * 14 * 14, 256),
          (512 * 14 * 14, 1024), (512 * 7 * 7, 512), (512 * 7 * 7, 2048)]


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


rows = []
for M, C in SHAPES:
    nrot = max(1, min(4, int(400e6 // (M * C * 2)) + 1))
    xs = [torch.randn(M, C, device=dev).bfloat16().requires_grad_(True) for _ in range(nrot)]
    g = torch.ones(C, device=dev, requires_grad=True)
    b = torch.zeros(C, device=dev, requires_grad=True)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    ys = [K.bn_act(x, g, b, rm, rv, relu=True) for x in xs]
    dys = [torch.randn(M, C, device=dev).bfloat16() for _ in range(nrot)]
    it = [0]

    def fwd():
        i = it[0] % nrot
        it[0] += 1
        K.bn_act(xs[i].detach(), g.detach(), b.detach(), rm, rv, relu=True)

    def bwd():
        i = it[0] % nrot
        it[0] += 1
        ys[i].backward(dys[i], retain_graph=True)

    t_f = timeit(fwd)
    t_b = timeit(bwd)
    byts = M * C * 2
    row = {"M": M, "C": C, "fwd_stats_apply_us": t_f, "fwd_GBs": 3 * byts / t_f / 1e3, "bwd_reduce_apply_us": t_b,
           "bwd_GBs": 5 * byts / t_b / 1e3, "bwd_frac_of_measured_hbm": 5 * byts / t_b / 1e3 / HBM,
           "fwd_frac_of_measured_hbm": 3 * byts / t_f / 1e3 / HBM}
    rows.append(row)
    print(json.dumps(row), flush=True)
if len(sys.argv) > 1:
    json.dump({"hbm_gbs": HBM, "rows": rows, "note": "fwd = stats pass (read x) + apply (read x, write y [+mask]); bwd = reduce (read dy, x) + "
               "apply (read dy, x, write dx): algorithmic bytes 3x / 5x the tensor"}, open(sys.argv[1], "w"), indent=1)
