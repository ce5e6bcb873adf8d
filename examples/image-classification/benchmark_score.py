#!/usr/bin/env python
"""Inference throughput of the model-zoo networks (reference: example/image-classification/benchmark_score.py, whose
V100 numbers are quoted in docs/faq/perf.md / BASELINE.md).  Forward only, synthetic data, device-timed.

    python examples/image-classification/benchmark_score.py --networks resnet50_v1b,resnet152_v1 --batch-sizes 32,128 --dtype bfloat16
"""
import argparse
import json
import logging
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mxb200 as mx  # noqa: E402
from mxb200 import autograd  # noqa: E402
from mxb200.gluon.model_zoo import vision  # noqa: E402

logging.basicConfig(level=logging.INFO)


def score(network, dev, batch_size, num_batches, dtype, image=224):
    fused = network in ("resnet50_v1b", "resnet18_v1b", "resnet34_v1b", "resnet101_v1b", "resnet152_v1b") and dev.device_type == "gpu" \
        and dtype == "bfloat16"
    if fused:
        layers = int(network.split("_")[0].replace("resnet", ""))
        net = vision.get_resnet("v1b", layers, ctx=dev, dtype="bfloat16", layout="NHWC")
        shape = (batch_size, image, image, 3)
    else:
        net = vision.get_model(network)
        net.initialize(ctx=dev)
        if dtype != "float32":
            net.cast(dtype)
        s = 299 if "inception" in network else image
        shape = (batch_size, 3, s, s)
    net.hybridize()
    x = mx.nd.random.uniform(-1, 1, shape=shape, ctx=dev).astype(dtype)
    with autograd.predict_mode():
        for _ in range(5):
            net(x).wait_to_read()
        mx.nd.waitall()
        if dev.device_type == "gpu":
            import torch
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(num_batches):
                y = net(x)
            e1.record()
            torch.cuda.synchronize()
            sec = e0.elapsed_time(e1) / 1e3
        else:
            tic = time.time()
            for _ in range(num_batches):
                y = net(x)
            y.wait_to_read()
            sec = time.time() - tic
    return num_batches * batch_size / sec


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--networks", default="alexnet,vgg16,inceptionv3,resnet50_v1,resnet50_v1b,resnet152_v1")
    ap.add_argument("--batch-sizes", default="1,32,128")
    ap.add_argument("--num-batches", type=int, default=20)
    ap.add_argument("--dtype", default="bfloat16")
    ap.add_argument("--gpu", type=int, default=0)
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    dev = mx.gpu(args.gpu) if mx.context.num_gpus() > 0 else mx.cpu()
    rows = []
    for net in args.networks.split(","):
        logging.info("network: %s  device: %s  dtype: %s", net, dev, args.dtype)
        for b in [int(v) for v in args.batch_sizes.split(",")]:
            speed = score(net, dev, b, args.num_batches, args.dtype if dev.device_type == "gpu" else "float32")
            logging.info("batch size %2d, dtype %s, images/sec: %f", b, args.dtype, speed)
            rows.append({"network": net, "batch": b, "dtype": args.dtype, "images_per_sec": speed})
    if args.json:
        with open(args.json, "w") as f:
            json.dump(rows, f, indent=1)
