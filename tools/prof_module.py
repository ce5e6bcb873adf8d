#!/usr/bin/env python
"""Host + device profile of the generic symbolic path: gluon model-zoo ResNet traced to a Symbol (NCHW), Module
forward/backward/update on one GPU with synthetic data -- what examples/image-classification/train_imagenet.py runs.
Prints img/s, the top host functions (cProfile, cumulative) and the top kernels (torch.profiler, device time).

    python tools/prof_module.py [--batch 128] [--layers 50] [--dtype bfloat16] [--steps 10]
"""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import mxb200 as mx  # noqa: E402
from mxb200.gluon.model_zoo import vision  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--layers", type=int, default=50)
    ap.add_argument("--dtype", default="bfloat16")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--image", type=int, default=224)
    a = ap.parse_args()
    ctx = mx.gpu(0) if torch.cuda.is_available() else mx.cpu()
    net = vision.get_model("resnet%d_v1" % a.layers, classes=1000)
    data = mx.sym.var("data")
    if a.dtype != "float32":
        data = mx.sym.Cast(data=data, dtype=a.dtype)
        net.cast(a.dtype)
    out = net(data)
    if a.dtype != "float32":
        out = mx.sym.Cast(data=out, dtype="float32")
    sym = mx.sym.SoftmaxOutput(out, name="softmax")
    it = mx.io.SyntheticDataIter(1000, (a.batch, 3, a.image, a.image), a.steps, dtype="float32")
    mod = mx.mod.Module(sym, context=[ctx])
    mod.bind(it.provide_data, it.provide_label)
    mod.init_params(mx.init.Xavier(rnd_type="gaussian", factor_type="in", magnitude=2))
    mod.init_optimizer(kvstore="device", optimizer="sgd",
                       optimizer_params={"learning_rate": 0.1, "momentum": 0.9, "wd": 1e-4, "multi_precision": a.dtype != "float32"})
    metric = mx.metric.create("accuracy")

    def sync():
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def epoch():
        it.reset()
        for batch in it:
            mod.forward_backward(batch)
            mod.update()
            mod.update_metric(metric, batch.label)
        sync()
    epoch()                                   # warm-up (cuDNN autotune, arenas)
    t0 = time.time()
    epoch()
    dt = time.time() - t0
    print("fused update: %s   %.1f img/s  (%.1f ms / batch of %d)" % (getattr(mod, "_fused", None) is not None,
                                                                      a.steps * a.batch / dt, 1e3 * dt / a.steps, a.batch))
    pr = cProfile.Profile()
    pr.enable()
    epoch()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(40)
    print("\n".join(l[:170] for l in s.getvalue().splitlines() if l.strip())[:9000])
    if torch.cuda.is_available():
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            epoch()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70)[:12000])


if __name__ == "__main__":
    main()
