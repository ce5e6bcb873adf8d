#!/usr/bin/env python
"""Per-shape timing of the implicit-GEMM tcgen05 convolution against cuDNN (through torch) for every 3x3 convolution of
ResNet-50 v1b at the benchmark batch: forward, data gradient, weight gradient.  CUDA events, 3 warm-up + 10 timed
iterations over rotating operand copies (so that the small late-stage tensors do not live in the 126 MB L2)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

SHAPES = [  # (h, w, cin, cout, stride, count in ResNet-50)
    (56, 56, 64, 64, 1, 3), (56, 56, 128, 128, 2, 1), (28, 28, 128, 128, 1, 3), (28, 28, 256, 256, 2, 1),
    (14, 14, 256, 256, 1, 5), (14, 14, 512, 512, 2, 1), (7, 7, 512, 512, 1, 2),
]


def timeit(fn, nrot, iters=10, warm=3):
    for i in range(warm):
        fn(i % nrot)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % nrot)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3       # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    from mxb200.ops import functional as K
    dev = torch.device("cuda", 0)
    rows = []
    tot = {"ours": 0.0, "cudnn": 0.0}
    for (h, w, c, o, s, cnt) in SHAPES:
        n = args.batch
        P, Q = (h + 2 - 3) // s + 1, (w + 2 - 3) // s + 1
        xbytes = n * h * w * c * 2
        nrot = max(1, min(6, int(300e6 // xbytes) + 1))
        xs = [torch.randn(n, h, w, c, device=dev).bfloat16() for _ in range(nrot)]
        dys = [torch.randn(n, P, Q, o, device=dev).bfloat16() for _ in range(nrot)]
        wk = (torch.randn(o, 3, 3, c, device=dev) * 0.05).bfloat16()
        wt = wk.permute(0, 3, 1, 2)
        dw = torch.zeros(o, 3, 3, c, device=dev)
        flops = 2.0 * n * P * Q * o * 9 * c
        r = {"shape": "n%d %dx%dx%d->%d s%d" % (n, h, w, c, o, s), "count": cnt, "gflop": flops / 1e9}
        r["fwd_us"] = timeit(lambda i: K.conv_fwd_raw(xs[i], wk, s, 1), nrot)
        r["dgrad_us"] = timeit(lambda i: K.conv_dgrad_raw(dys[i], wk, xs[i].shape, s, 1), nrot)
        r["wgrad_us"] = timeit(lambda i: K.conv_wgrad_raw(dys[i], xs[i], wk.shape, s, 1, out=dw), nrot)
        xc = [x.permute(0, 3, 1, 2) for x in xs]
        dyc = [d.permute(0, 3, 1, 2) for d in dys]
        r["cudnn_fwd_us"] = timeit(lambda i: F.conv2d(xc[i], wt, None, s, 1), nrot)
        r["cudnn_dgrad_us"] = timeit(lambda i: torch.ops.aten.convolution_backward(
            dyc[i], xc[i], wt, None, [s, s], [1, 1], [1, 1], False, [0, 0], 1, [True, False, False]), nrot)
        r["cudnn_wgrad_us"] = timeit(lambda i: torch.ops.aten.convolution_backward(
            dyc[i], xc[i], wt, None, [s, s], [1, 1], [1, 1], False, [0, 0], 1, [False, True, False]), nrot)
        for k in ("fwd", "dgrad", "wgrad"):
            r[k + "_tflops"] = flops / r[k + "_us"] / 1e6
            r[k + "_vs_cudnn"] = r["cudnn_" + k + "_us"] / r[k + "_us"]
            tot["ours"] += cnt * r[k + "_us"]
            tot["cudnn"] += cnt * r["cudnn_" + k + "_us"]
        rows.append(r)
        print(json.dumps(r), flush=True)
    summary = {"total_us_per_step_ours": tot["ours"], "total_us_per_step_cudnn": tot["cudnn"], "batch": args.batch}
    print(json.dumps(summary))
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"rows": rows, "summary": summary}, f, indent=1)


if __name__ == "__main__":
    main()
