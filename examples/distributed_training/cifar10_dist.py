#!/usr/bin/env python
"""Distributed data-parallel training of a small ResNet on CIFAR-10 with Gluon + a distributed kvstore.

Counterpart of the reference's ``example/distributed_training/cifar10_dist.py`` (same recipe: every worker trains on its
own shard chosen with a SplitSampler, gradients are aggregated by the kvstore, the Trainer applies the update), written
for this framework's two process models:

  * parameter-server job (``--kv-store dist_sync`` / ``dist_device_sync`` / ``dist_async``), started with our launcher
        python tools/launch.py -n 2 -s 2 --launcher local python examples/distributed_training/cifar10_dist.py
  * one process per GPU over NVLink (``--kv-store device`` under torchrun): gradient exchange + optimizer are ONE fused
    kernel per step (csrc/cuda/allreduce.cu)
        python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/distributed_training/cifar10_dist.py --kv-store device

There is no network on the build machines, so when the CIFAR-10 archive is not under ``--data-dir`` a synthetic dataset
of the same shape is used (``--synthetic`` forces it).
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mxb200 as mx  # noqa: E402
from mxb200 import autograd, gluon  # noqa: E402
from mxb200.gluon.data.sampler import SplitSampler  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kv-store", default="dist_sync")
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--batch-size", type=int, default=64, help="per device")
    ap.add_argument("--gpus", type=int, default=-1, help="devices of THIS process (-1: all visible, 0: CPU)")
    ap.add_argument("--lr", type=float, default=0.1)
    ap.add_argument("--model", default="resnet18_v1")
    ap.add_argument("--data-dir", default=os.path.expanduser("~/.mxnet/datasets/cifar10"))
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--samples", type=int, default=2048, help="size of the synthetic dataset")
    return ap.parse_args()


class SyntheticCifar(gluon.data.Dataset):
    """Class-dependent blobs: learnable, so the accuracy printed per epoch means something."""

    def __init__(self, n, seed=0):
        rng = np.random.default_rng(seed)
        self.label = rng.integers(0, 10, n).astype(np.float32)
        centers = rng.normal(0, 1, (10, 3, 1, 1)).astype(np.float32)
        self.data = (rng.normal(0, 0.5, (n, 3, 32, 32)).astype(np.float32) + centers[self.label.astype(int)])

    def __len__(self):
        return len(self.label)

    def __getitem__(self, i):
        return self.data[i], self.label[i]


def transform(data, label):
    # HWC uint8 -> CHW float in [0, 1]
    return mx.nd.transpose(data.astype(np.float32), (2, 0, 1)) / 255, label.astype(np.float32)


def main():
    args = parse()
    store = mx.kv.create(args.kv_store)
    import torch
    ngpu = torch.cuda.device_count() if args.gpus < 0 else args.gpus
    if "LOCAL_RANK" in os.environ and ngpu > 0:          # torchrun: this process owns one GPU
        ctx = [mx.gpu(int(os.environ["LOCAL_RANK"]))]
    else:
        ctx = [mx.gpu(i) for i in range(ngpu)] if ngpu > 0 else [mx.cpu()]
    batch = args.batch_size * len(ctx)

    have_real = os.path.exists(os.path.join(args.data_dir, "cifar-10-binary.tar.gz")) or os.path.isdir(
        os.path.join(args.data_dir, "cifar-10-batches-bin"))
    if args.synthetic or not have_real:
        train_set, test_set = SyntheticCifar(args.samples, 0), SyntheticCifar(args.samples // 4, 1)
    else:
        train_set = gluon.data.vision.CIFAR10(args.data_dir, train=True, transform=transform)
        test_set = gluon.data.vision.CIFAR10(args.data_dir, train=False, transform=transform)
    # every worker samples only from its own part of the training set
    train_data = gluon.data.DataLoader(train_set, batch, sampler=SplitSampler(len(train_set), store.num_workers, store.rank),
                                       last_batch="discard")
    test_data = gluon.data.DataLoader(test_set, batch, shuffle=False)

    net = getattr(gluon.model_zoo.vision, args.model)(classes=10, thumbnail=True) if "resnet" in args.model else \
        getattr(gluon.model_zoo.vision, args.model)(classes=10)
    net.initialize(mx.init.Xavier(magnitude=2), ctx=ctx)
    net.hybridize()
    loss_fn = gluon.loss.SoftmaxCrossEntropyLoss()
    trainer = gluon.Trainer(net.collect_params(), "sgd", {"learning_rate": args.lr, "momentum": 0.9, "wd": 1e-4}, kvstore=store)

    def evaluate():
        acc = mx.metric.Accuracy()
        for data, label in test_data:
            xs = gluon.utils.split_and_load(data, ctx, even_split=False)
            ys = gluon.utils.split_and_load(label, ctx, even_split=False)
            acc.update(ys, [net(x) for x in xs])
        return acc.get()[1]

    for epoch in range(args.epochs):
        tic, seen = time.time(), 0
        for data, label in train_data:
            xs = gluon.utils.split_and_load(data, ctx)
            ys = gluon.utils.split_and_load(label, ctx)
            with autograd.record():
                losses = [loss_fn(net(x), y) for x, y in zip(xs, ys)]
            for l in losses:
                l.backward()
            trainer.step(batch)          # kvstore aggregates across devices AND workers
            seen += batch
        mx.nd.waitall()
        dt = time.time() - tic
        print("worker %d/%d epoch %d: %.0f samples/s, validation accuracy %.4f" % (store.rank, store.num_workers, epoch, seen / dt,
                                                                                   evaluate()), flush=True)


if __name__ == "__main__":
    main()
