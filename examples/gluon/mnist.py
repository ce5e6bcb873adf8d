#!/usr/bin/env python
"""Gluon image classification on MNIST-shaped data: HybridSequential MLP / small conv net, DataLoader, Trainer, metric,
save / load of parameters (reference: example/gluon/mnist.py, image_classification.py).

    python examples/gluon/mnist.py [--gpu 0] [--epochs 3] [--net cnn]

When the MNIST files are not under --data-dir a synthetic 10-class dataset of the same shape is used (no network here)."""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mxb200 as mx  # noqa: E402
from mxb200 import autograd, gluon  # noqa: E402
from mxb200.gluon import nn  # noqa: E402


class SyntheticDigits(gluon.data.Dataset):
    """Ten fixed random 28x28 prototypes + noise: separable but not trivially so."""

    def __init__(self, n, seed):
        rng = np.random.default_rng(seed)
        protos = np.random.default_rng(1234).normal(0, 1, (10, 1, 28, 28)).astype(np.float32)
        self.y = rng.integers(0, 10, n).astype(np.float32)
        self.x = protos[self.y.astype(int)] + rng.normal(0, 2.0, (n, 1, 28, 28)).astype(np.float32)

    def __len__(self):
        return len(self.y)

    def __getitem__(self, i):
        return self.x[i], self.y[i]


def build(kind):
    net = nn.HybridSequential()
    with net.name_scope():
        if kind == "cnn":
            net.add(nn.Conv2D(16, 3, padding=1, activation="relu"), nn.MaxPool2D(2), nn.Conv2D(32, 3, padding=1, activation="relu"),
                    nn.MaxPool2D(2), nn.Flatten(), nn.Dense(64, activation="relu"), nn.Dense(10))
        else:
            net.add(nn.Flatten(), nn.Dense(128, activation="relu"), nn.Dense(64, activation="relu"), nn.Dense(10))
    return net


def evaluate(net, data, ctx):
    acc = mx.metric.Accuracy()
    for x, y in data:
        acc.update([y.as_in_context(ctx)], [net(x.as_in_context(ctx))])
    return acc.get()[1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu", type=int, default=-1)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--batch-size", type=int, default=100)
    ap.add_argument("--net", default="mlp", choices=["mlp", "cnn"])
    ap.add_argument("--lr", type=float, default=None, help="default 0.1 (mlp) / 0.02 (cnn)")
    ap.add_argument("--samples", type=int, default=6000)
    ap.add_argument("--data-dir", default=os.path.expanduser("~/.mxnet/datasets/mnist"))
    args = ap.parse_args()
    if args.lr is None:
        args.lr = 0.1 if args.net == "mlp" else 0.02
    ctx = mx.gpu(args.gpu) if args.gpu >= 0 else mx.cpu()
    if os.path.exists(os.path.join(args.data_dir, "train-images-idx3-ubyte.gz")):
        tf = lambda d, l: (d.astype(np.float32).transpose((2, 0, 1)) / 255, np.float32(l))   # noqa: E731
        train_set = gluon.data.vision.MNIST(args.data_dir, train=True, transform=tf)
        test_set = gluon.data.vision.MNIST(args.data_dir, train=False, transform=tf)
    else:
        train_set, test_set = SyntheticDigits(args.samples, 0), SyntheticDigits(args.samples // 5, 1)
    train = gluon.data.DataLoader(train_set, args.batch_size, shuffle=True, last_batch="discard")
    test = gluon.data.DataLoader(test_set, args.batch_size)
    net = build(args.net)
    net.initialize(mx.init.Xavier(), ctx=ctx)
    net.hybridize()
    trainer = gluon.Trainer(net.collect_params(), "sgd", {"learning_rate": args.lr, "momentum": 0.9})
    loss_fn = gluon.loss.SoftmaxCrossEntropyLoss()
    for epoch in range(args.epochs):
        tic, metric = time.time(), mx.metric.Accuracy()
        for x, y in train:
            x, y = x.as_in_context(ctx), y.as_in_context(ctx)
            with autograd.record():
                out = net(x)
                loss = loss_fn(out, y)
            loss.backward()
            trainer.step(x.shape[0])
            metric.update([y], [out])
        print("epoch %d: train accuracy %.4f, test accuracy %.4f, %.1f s" % (epoch, metric.get()[1], evaluate(net, test, ctx),
                                                                          time.time() - tic), flush=True)
    f = os.path.join(tempfile.gettempdir(), "gluon_mnist.params")
    net.save_parameters(f)
    net2 = build(args.net)
    net2.load_parameters(f, ctx=ctx)
    acc = evaluate(net2, test, ctx)
    print("reloaded parameters: test accuracy %.4f" % acc)
    return acc


if __name__ == "__main__":
    main()
