#!/usr/bin/env python
"""Profile the forward / backward passes of a bound symbolic executor (reference: example/profiler/profiler_executor.py).

    python examples/profiler/profiler_executor.py [--gpu 0] [--iters 10]"""
import argparse
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mxb200 as mx  # noqa: E402
from mxb200 import profiler  # noqa: E402


def mlp(hidden=256, classes=10):
    data = mx.sym.Variable("data")
    h = mx.sym.Activation(mx.sym.FullyConnected(data, num_hidden=hidden, name="fc1"), act_type="relu")
    h = mx.sym.Activation(mx.sym.FullyConnected(h, num_hidden=hidden, name="fc2"), act_type="relu")
    return mx.sym.SoftmaxOutput(mx.sym.FullyConnected(h, num_hidden=classes, name="fc3"), name="softmax")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu", type=int, default=-1)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--file", default=os.path.join(tempfile.gettempdir(), "profile_executor.json"))
    args = ap.parse_args()
    ctx = mx.gpu(args.gpu) if args.gpu >= 0 else mx.cpu()
    ex = mlp().simple_bind(ctx, data=(args.batch_size, 784), grad_req="write")
    for name, arr in ex.arg_dict.items():
        if name not in ("data", "softmax_label"):
            arr[:] = mx.nd.random.normal(0, 0.05, shape=arr.shape, ctx=ctx)
    ex.arg_dict["data"][:] = mx.nd.random.uniform(0, 1, shape=(args.batch_size, 784), ctx=ctx)
    ex.arg_dict["softmax_label"][:] = mx.nd.array([i % 10 for i in range(args.batch_size)], ctx=ctx)
    ex.forward(is_train=True)
    ex.backward()
    mx.nd.waitall()
    profiler.set_config(profile_symbolic=True, filename=args.file)
    profiler.set_state("run")
    for _ in range(args.iters):
        with profiler.scope("forward", cat="executor", device=True):
            ex.forward(is_train=True)
        with profiler.scope("backward", cat="executor", device=True):
            ex.backward()
    mx.nd.waitall()
    profiler.set_state("stop")
    print(profiler.dumps())
    profiler.dump()
    ev = json.load(open(args.file))["traceEvents"]
    print("%d trace events written to %s" % (len(ev), args.file))
    return ev


if __name__ == "__main__":
    main()
