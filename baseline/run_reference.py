#!/usr/bin/env python
"""Reference arm driver: runs the UNMODIFIED reference (MXNet 1.3.1 fork built by baseline/build_ref.sh into
baseline/_ref) through its own public API and stock code path, for the headline metric and config.

  network : example/image-classification/symbols/resnet.py (the network of the reference's own benchmark table,
            docs/faq/perf.md:207-216: ``train_imagenet.py --benchmark 1 --network resnet --num-layers 50``)
  loop    : mx.mod.Module bind / init_params / init_optimizer(kvstore='device') / forward_backward / update -- what
            Module.fit does per batch (python/mxnet/module/base_module.py:520-560), N GPUs driven by ONE process (the
            reference's process model, executor_group.py:143-305)
  data    : synthetic, one pinned-host batch re-used for ``value`` (common/data.py:98-130 SyntheticDataIter semantics); for ``e2e`` every
            step feeds a host numpy batch (uint8 -> float on host like ImageRecordIter would deliver floats) and reads the
            softmax output of the step back to the host.

No file of this repository's framework is imported here.  Timing is host wall-clock around mx.nd.waitall() (the reference has
no CUDA-event API in python); a full device drain brackets both sides.
"""
import argparse
import json
import os
import sys
import time

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=5)
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--image", type=int, default=224)
ap.add_argument("--classes", type=int, default=1000)
ap.add_argument("--layers", type=int, default=50)
ap.add_argument("--dtype", default="float16")
args = ap.parse_args()

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_ref"))
sys.path.insert(0, os.path.join(HERE, "_ref", "example", "image-classification"))
os.environ.setdefault("MXNET_CUDNN_AUTOTUNE_DEFAULT", "0")

import numpy as np  # noqa: E402
import mxnet as mx  # noqa: E402
from importlib import import_module  # noqa: E402

N = args.gpus
B = args.batch * N
S = args.image
ctxs = [mx.gpu(i) for i in range(N)]
sym = import_module("symbols.resnet").get_symbol(num_classes=args.classes, num_layers=args.layers,
                                                 image_shape="3,%d,%d" % (S, S), dtype=args.dtype)
mod = mx.mod.Module(sym, context=ctxs, data_names=["data"], label_names=["softmax_label"])
mod.bind(data_shapes=[("data", (B, 3, S, S))], label_shapes=[("softmax_label", (B,))], for_training=True)
mod.init_params(initializer=mx.init.Xavier(rnd_type="gaussian", factor_type="in", magnitude=2))
kv = mx.kv.create("device") if N > 1 else None
mod.init_optimizer(kvstore=kv if kv is not None else "local", optimizer="nag",
                   optimizer_params={"learning_rate": 0.1, "momentum": 0.9, "wd": 1e-4,
                                     "multi_precision": args.dtype == "float16", "rescale_grad": 1.0 / B})

rng = np.random.RandomState(1234)
host_img = rng.randint(0, 256, size=(B, 3, S, S)).astype(np.uint8)
host_lab = rng.randint(0, args.classes, size=(B,)).astype(np.float32)
# the synthetic batch (SyntheticDataIter keeps one pinned-host NDArray and returns it every iteration)
pinned = mx.Context("cpu_pinned", 0)      # where the reference's SyntheticDataIter keeps its batch (common/data.py:110-111)
res = mx.io.DataBatch(data=[mx.nd.array(host_img.astype(np.float32), ctx=pinned)],
                      label=[mx.nd.array(host_lab, ctx=pinned)])


def step(batch):
    mod.forward_backward(batch)
    mod.update()


W = max(3, args.warmup)
for _ in range(W):
    step(res)
mx.nd.waitall()
t0 = time.time()
for _ in range(args.steps):
    step(res)
mx.nd.waitall()
ms = (time.time() - t0) * 1e3
value = B * args.steps / (ms / 1e3)

mean = np.array([123.68, 116.28, 103.53], dtype=np.float32).reshape(1, 3, 1, 1)
istd = 1.0 / np.array([58.395, 57.12, 57.375], dtype=np.float32).reshape(1, 3, 1, 1)
last = [0.0]


def step_e2e():
    x = mx.nd.array(host_img, ctx=mx.cpu(), dtype=np.uint8)               # host batch (uint8, like decoded records)
    xf = (x.astype(np.float32) - mx.nd.array(mean)) * mx.nd.array(istd)  # ImageRecordIter-style host normalisation
    b = mx.io.DataBatch(data=[xf], label=[mx.nd.array(host_lab)])
    step(b)
    out = mod.get_outputs()[0]
    last[0] = float(out[0:1].asnumpy().ravel()[0])                        # device->host read of this step's result


for _ in range(2):
    step_e2e()
mx.nd.waitall()
t0 = time.time()
for _ in range(args.steps):
    step_e2e()
mx.nd.waitall()
ms2 = (time.time() - t0) * 1e3

print(json.dumps({
    "metric": "ResNet-50 training throughput", "value": value, "unit": "images/sec", "n_gpus": N, "steps": args.steps,
    "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
    "vs_baseline": value / 363.69, "dtype": args.dtype, "impl": "reference",
    "data": "synthetic (random images / labels, random-init weights)",
    "config": {"model": "resnet%d (example/image-classification/symbols/resnet.py)" % args.layers, "global_batch": B,
               "per_gpu_batch": args.batch, "image": [S, S, 3], "classes": args.classes, "parallelism": "dp%d" % N,
               "optimizer": "NAG mom=0.9 wd=1e-4" + (" multi_precision" if args.dtype == "float16" else ""),
               "kvstore": "device (one process, %d GPUs)" % N,
               "build": "MXNet 1.3.1 fork from /root/reference, USE_CUDA=1 USE_CUDNN=0 (cuDNN-7 API absent), sm_100",
               "timing": "host wall clock around mx.nd.waitall()"},
    "gpu_launches": 0,
    "e2e": {"value": B * args.steps / (ms2 / 1e3), "unit": "images/sec", "ms_per_step": ms2 / args.steps,
            "h2d_bytes_per_step": int(host_img.size * 4 + host_lab.size * 4), "d2h_bytes_per_step": 4},
}))
