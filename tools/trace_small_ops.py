#!/usr/bin/env python
"""Where do the small library kernels of a training step come from?

Runs the bench's ResNet-50 step (gluon model + fused Trainer) a few times, then one more step under a
TorchDispatchMode that records every aten op touching a CUDA tensor together with the innermost python frames of
this repo.  Prints / writes a table (op, numel bucket, call site) -> count, so stray fills / adds / casts of the step
can be traced back to a line of python (the per-kernel ncu table only has kernel names).

    python tools/trace_small_ops.py [--batch 64] [--out gpurun_out/small_ops.txt]
"""
import argparse
import collections
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
from torch.utils._python_dispatch import TorchDispatchMode  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Recorder(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.rows = collections.Counter()

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        out = func(*args, **(kwargs or {}))
        name = str(func)
        if any(s in name for s in ("view", "detach", "as_strided", "alias", "empty", "_unsafe_view", "reshape", "t.default",
                                   "permute", "slice", "select", "expand", "unsqueeze", "squeeze", "transpose",
                                   "is_", "size", "stride", "storage_offset", "numel", "dim")):
            return out
        tens = [a for a in list(args) + list((kwargs or {}).values()) if isinstance(a, torch.Tensor)]
        if isinstance(out, torch.Tensor):
            tens.append(out)
        if not any(t.is_cuda for t in tens):
            return out
        n = max((t.numel() for t in tens), default=0)
        bucket = "<=4K" if n <= 4096 else "<=1M" if n <= (1 << 20) else ">1M"
        frames = [f for f in traceback.extract_stack()
                  if f.filename.startswith(ROOT) and "trace_small_ops" not in f.filename]
        site = " < ".join("%s:%d" % (os.path.relpath(f.filename, ROOT), f.lineno) for f in frames[-3:][::-1]) or "(autograd engine)"
        dts = ",".join(sorted({str(t.dtype).replace("torch.", "") for t in tens}))
        self.rows[(name, dts, bucket, site)] += 1
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import mxb200 as mx
    from mxb200 import gluon, autograd
    from mxb200.gluon.model_zoo import vision
    ctx = mx.gpu(0)
    net = vision.get_resnet("v1b", 50, classes=1000, ctx=ctx, dtype="bfloat16", layout="NHWC")
    net.hybridize()
    trainer = gluon.Trainer(net.collect_params(), "nag", {"learning_rate": 0.1, "momentum": 0.9, "wd": 1e-4,
                                                          "multi_precision": True}, kvstore=mx.kv.create("device"))
    loss_fn = gluon.loss.SoftmaxCrossEntropyLoss()
    x = mx.nd.from_torch(torch.randn(a.batch, 224, 224, 3, device="cuda").to(torch.bfloat16), ctx)
    y = mx.nd.from_torch(torch.randint(0, 1000, (a.batch,), device="cuda"), ctx)

    def step():
        with autograd.record():
            loss = loss_fn(net(x), y)
        loss.backward()
        trainer.step(a.batch)
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    rec = Recorder()
    with rec:
        step()
    torch.cuda.synchronize()
    lines = ["%5d  %-34s %-18s %-5s %s" % (c, k[0], k[1], k[2], k[3]) for k, c in rec.rows.most_common()]
    text = "aten ops on CUDA tensors in ONE training step (count, op, dtypes, size, call site)\n" + "\n".join(lines)
    print(text)
    if a.out:
        with open(a.out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
