#!/usr/bin/env python
"""Diagnostic: ResNet-18 thumbnail, fused arena Trainer vs per-parameter updaters, per-parameter error after k steps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mxb200 as mx  # noqa: E402
from mxb200 import autograd, gluon  # noqa: E402
from mxb200.gluon.model_zoo import vision  # noqa: E402

ctx = mx.gpu(0)
torch.manual_seed(0)
x = mx.nd.from_torch(torch.randn(8, 32, 32, 3, device="cuda").bfloat16(), ctx)
y = mx.nd.from_torch(torch.randint(0, 10, (8,), device="cuda"), ctx)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
res = {}
for fused in ("1", "0", "1b"):
    os.environ["MXB200_FUSED_TRAINER"] = fused[0]
    net = vision.get_resnet("v1b", 18, classes=10, thumbnail=True, ctx=ctx, dtype="bfloat16", layout="NHWC", seed=3)
    tr = gluon.Trainer(net.collect_params(), "nag", {"learning_rate": 0.05, "momentum": 0.9, "wd": 1e-4,
                                                     "multi_precision": True})
    L = gluon.loss.SoftmaxCrossEntropyLoss()
    snaps, losses, grads = [], [], []
    for it in range(steps):
        with autograd.record():
            l = L(net(x), y)
        l.backward()
        if it == 0:
            grads = [(n, p.grad().as_torch().float().clone()) for n, p in net.collect_params().items() if p.grad_req != "null"]
        tr.step(8)
        losses.append(float(l.as_torch().float().mean()))
        snaps.append([(n, p.data().as_torch().float().clone()) for n, p in net.collect_params().items()])
    res[fused] = (snaps, losses, grads)
print("losses fused  ", res["1"][1])
print("losses unfused", res["0"][1])
print("losses fused#2", res["1b"][1])


def rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-6)).item()


print("--- first-step gradients (fused vs unfused), then (fused vs fused#2 = run-to-run noise)")
for (n, ga), (_, gb), (_, gc) in zip(res["1"][2], res["0"][2], res["1b"][2]):
    e1, e2 = rel(ga, gb), rel(ga, gc)
    if e1 > 2e-2 or e2 > 1e-3:
        print("  grad %-45s %s fused-vs-unfused %.4f   run-to-run %.4f" % (n, tuple(ga.shape), e1, e2))
for k in range(steps):
    worst = sorted(((rel(a, b), n) for (n, a), (_, b) in zip(res["1"][0][k], res["0"][0][k])), reverse=True)[:6]
    noise = sorted(((rel(a, b), n) for (n, a), (_, b) in zip(res["1"][0][k], res["1b"][0][k])), reverse=True)[:3]
    print("step %d worst fused-vs-unfused:" % (k + 1), ["%s %.4f" % (n, e) for e, n in worst])
    print("        run-to-run (fused twice):", ["%s %.4f" % (n, e) for e, n in noise])
