#!/usr/bin/env python
"""Training-dynamics check of the generic NCHW path: the same symbolic bottleneck network, same initial weights and
batches, trained (a) on the library operators + per-key updater and (b) on the operator fast paths + graph pass + fused
Module.update.  Prints the per-step loss of both; they must track each other to bf16 noise."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mxb200 as mx  # noqa: E402


def block(data, name, nf, stride, down):
    c1 = mx.sym.Convolution(data, kernel=(1, 1), stride=(stride, stride), num_filter=nf // 4, no_bias=True, name=name + "_c1")
    a1 = mx.sym.Activation(mx.sym.BatchNorm(c1, fix_gamma=False, eps=1e-5, name=name + "_bn1"), act_type="relu")
    c2 = mx.sym.Convolution(a1, kernel=(3, 3), pad=(1, 1), num_filter=nf // 4, no_bias=True, name=name + "_c2")
    a2 = mx.sym.Activation(mx.sym.BatchNorm(c2, fix_gamma=False, eps=1e-5, name=name + "_bn2"), act_type="relu")
    c3 = mx.sym.Convolution(a2, kernel=(1, 1), num_filter=nf, no_bias=True, name=name + "_c3")
    b3 = mx.sym.BatchNorm(c3, fix_gamma=False, eps=1e-5, name=name + "_bn3")
    sc = data
    if down:
        sc = mx.sym.BatchNorm(mx.sym.Convolution(data, kernel=(1, 1), stride=(stride, stride), num_filter=nf, no_bias=True,
                                                 name=name + "_sc"), fix_gamma=False, eps=1e-5, name=name + "_scbn")
    return mx.sym.Activation(b3 + sc, act_type="relu")


def build():
    d = mx.sym.Cast(mx.sym.var("data"), dtype="bfloat16")
    n = mx.sym.Convolution(d, kernel=(3, 3), pad=(1, 1), num_filter=64, no_bias=True, name="stem")
    n = mx.sym.Activation(mx.sym.BatchNorm(n, fix_gamma=False, eps=1e-5, name="stem_bn"), act_type="relu")
    n = mx.sym.Pooling(n, kernel=(3, 3), stride=(2, 2), pad=(1, 1), pool_type="max")
    n = block(n, "u1", 256, 1, True)
    n = block(n, "u2", 256, 1, False)
    n = block(n, "u3", 512, 2, True)
    n = mx.sym.Pooling(n, global_pool=True, pool_type="avg", kernel=(1, 1))
    n = mx.sym.FullyConnected(mx.sym.flatten(n), num_hidden=16, name="fc")
    return mx.sym.SoftmaxOutput(mx.sym.Cast(n, dtype="float32"), name="softmax")


def run(fast, steps, init):
    for k in ("MXB200_OP_FASTPATH", "MXB200_GRAPH_FUSION", "MXB200_FUSED_MODULE"):
        os.environ[k] = fast
    rs = np.random.RandomState(1)
    x = rs.randn(256, 3, 32, 32).astype("f")
    y = ((x[:, 0, :16].mean(axis=(1, 2)) > 0).astype("i") * 8 + (x[:, 1, 16:].mean(axis=(1, 2)) > 0).astype("i") * 4
         + (x[:, 2].mean(axis=(1, 2)) > 0).astype("i")).astype("f")
    it = mx.io.NDArrayIter(x, y, batch_size=64)
    mod = mx.mod.Module(build(), context=[mx.gpu(0)])
    mod.bind(it.provide_data, it.provide_label)
    if init is None:
        mx.random.seed(5)
        mod.init_params(mx.init.Xavier(rnd_type="gaussian", factor_type="in", magnitude=2))
        init = tuple({k: v.copy() for k, v in d.items()} for d in mod.get_params())
    else:
        mod.init_params(arg_params=init[0], aux_params=init[1])
    mod.init_optimizer(kvstore="device", optimizer="sgd",
                       optimizer_params={"learning_rate": 0.05, "momentum": 0.9, "wd": 1e-4, "multi_precision": True})
    losses = []
    step = 0
    while step < steps:
        it.reset()
        for batch in it:
            mod.forward_backward(batch)
            p = mod.get_outputs()[0].asnumpy()
            lab = batch.label[0].asnumpy().astype(int)
            losses.append(float(-np.log(np.maximum(p[np.arange(len(lab)), lab], 1e-12)).mean()))
            mod.update()
            step += 1
            if step >= steps:
                break
    return losses, init, getattr(mod, "_fused", None) is not None


if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    l0, init, f0 = run("0", steps, None)
    l1, _, f1 = run("1", steps, init)
    print("library ops + per-key update (fused=%s) vs fast paths + graph pass + fused update (fused=%s)" % (f0, f1))
    for i in range(steps):
        print("step %2d  loss %.4f  %.4f" % (i, l0[i], l1[i]))
    a, b = np.array(l0), np.array(l1)
    print("mean |diff| first 5 steps %.4f ; mean loss last 10 steps %.4f vs %.4f" % (np.abs(a - b)[:5].mean(), a[-10:].mean(), b[-10:].mean()))
