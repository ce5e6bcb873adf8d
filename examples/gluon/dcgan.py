#!/usr/bin/env python
"""A small DCGAN with two Gluon trainers (reference: example/gluon/dcgan.py).

The training loop is the classic one: the discriminator step back-propagates through D twice (real and fake batch), the
generator step back-propagates through D into G -- which also deposits gradients in D's parameters that must NOT leak
into D's next update (`grad_req='write'`: every backward pass overwrites).  Synthetic "images": 16x16 Gaussian blobs.

    python examples/gluon/dcgan.py [--gpu 0] [--iters 200]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mxb200 as mx  # noqa: E402
from mxb200 import autograd, gluon  # noqa: E402
from mxb200.gluon import nn  # noqa: E402


def real_batch(n, rng):
    yy, xx = np.mgrid[0:16, 0:16].astype(np.float32)
    cx, cy = rng.uniform(4, 12, (n, 1, 1)), rng.uniform(4, 12, (n, 1, 1))
    img = np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / 8.0) * 2 - 1
    return img[:, None].astype(np.float32)


def generator(nz, ngf=32):
    g = nn.HybridSequential()
    with g.name_scope():
        g.add(nn.Conv2DTranspose(ngf * 2, 4, 1, 0, use_bias=False), nn.BatchNorm(), nn.Activation("relu"),       # 4x4
              nn.Conv2DTranspose(ngf, 4, 2, 1, use_bias=False), nn.BatchNorm(), nn.Activation("relu"),           # 8x8
              nn.Conv2DTranspose(1, 4, 2, 1, use_bias=False), nn.Activation("tanh"))                             # 16x16
    return g


def discriminator(ndf=32):
    d = nn.HybridSequential()
    with d.name_scope():
        d.add(nn.Conv2D(ndf, 4, 2, 1, use_bias=False), nn.LeakyReLU(0.2),                                        # 8x8
              nn.Conv2D(ndf * 2, 4, 2, 1, use_bias=False), nn.BatchNorm(), nn.LeakyReLU(0.2),                    # 4x4
              nn.Conv2D(1, 4, 1, 0, use_bias=False))                                                             # 1x1
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu", type=int, default=-1)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--nz", type=int, default=16)
    ap.add_argument("--lr", type=float, default=2e-4)
    args = ap.parse_args()
    ctx = mx.gpu(args.gpu) if args.gpu >= 0 else mx.cpu()
    rng = np.random.default_rng(0)
    netG, netD = generator(args.nz), discriminator()
    netG.initialize(mx.init.Normal(0.02), ctx=ctx)
    netD.initialize(mx.init.Normal(0.02), ctx=ctx)
    trainerG = gluon.Trainer(netG.collect_params(), "adam", {"learning_rate": args.lr, "beta1": 0.5})
    trainerD = gluon.Trainer(netD.collect_params(), "adam", {"learning_rate": args.lr, "beta1": 0.5})
    loss = gluon.loss.SigmoidBinaryCrossEntropyLoss()
    ones = mx.nd.ones((args.batch_size,), ctx=ctx)
    zeros = mx.nd.zeros((args.batch_size,), ctx=ctx)
    tic = time.time()
    for it in range(args.iters):
        real = mx.nd.array(real_batch(args.batch_size, rng), ctx=ctx)
        noise = mx.nd.random.normal(0, 1, shape=(args.batch_size, args.nz, 1, 1), ctx=ctx)
        # (1) D: maximise log D(x) + log(1 - D(G(z)))
        with autograd.record():
            errD_real = loss(netD(real).reshape((-1,)), ones)
            fake = netG(noise)
            errD_fake = loss(netD(fake.detach()).reshape((-1,)), zeros)
            errD = errD_real + errD_fake
        errD.backward()
        trainerD.step(args.batch_size)
        # (2) G: maximise log D(G(z)); this backward also writes gradients into D's parameters
        with autograd.record():
            errG = loss(netD(fake).reshape((-1,)), ones)
        errG.backward()
        trainerG.step(args.batch_size)
        if it % 50 == 0 or it == args.iters - 1:
            print("iter %4d  loss_D %.4f  loss_G %.4f  (%.1f it/s)" % (it, float(errD.mean().asscalar()), float(errG.mean().asscalar()),
                                                                        (it + 1) / (time.time() - tic)), flush=True)
    sample = netG(mx.nd.random.normal(0, 1, shape=(8, args.nz, 1, 1), ctx=ctx)).asnumpy()
    print("generated batch: shape %s, range [%.2f, %.2f], finite %s" % (sample.shape, sample.min(), sample.max(), bool(np.isfinite(sample).all())))
    return sample


if __name__ == "__main__":
    main()
