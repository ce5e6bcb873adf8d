#!/usr/bin/env python
"""Profile a matrix-multiply loop: operator scopes (host time + device time from CUDA events on a GPU), engine operations
(an asynchronous checkpoint write pushed while the loop runs) and the aggregate table (reference:
example/profiler/profiler_matmul.py, python/mxnet/profiler.py; `dump()` writes Chrome-trace JSON for chrome://tracing).

    python examples/profiler/profiler_matmul.py [--gpu 0] [--iters 20] [--file profile_matmul.json]"""
import argparse
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mxb200 as mx  # noqa: E402
from mxb200 import profiler  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu", type=int, default=-1)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--file", default=os.path.join(tempfile.gettempdir(), "profile_matmul.json"))
    args = ap.parse_args()
    ctx = mx.gpu(args.gpu) if args.gpu >= 0 else mx.cpu()
    a = mx.nd.random.uniform(-1, 1, shape=(args.size, args.size), ctx=ctx)
    b = mx.nd.random.uniform(-1, 1, shape=(args.size, args.size), ctx=ctx)
    profiler.set_config(profile_all=True, aggregate_stats=True, filename=args.file)
    for _ in range(30):                                  # warm-up outside the profile (library thread pools settle)
        mx.nd.dot(a, b).wait_to_read()
    profiler.set_state("run")
    ckpt = os.path.join(tempfile.gettempdir(), "profiler_matmul_ckpt.params")
    for i in range(args.iters):
        with profiler.scope("dot_%dx%d" % (args.size, args.size), cat="operator", device=True):
            c = mx.nd.dot(a, b)
        if i == args.iters // 2:
            mx.nd.save(ckpt, {"c": c}, async_write=True)  # host work on the dependency engine, shows up as cat "engine"
    c.wait_to_read()
    mx.nd.waitall()
    profiler.set_state("stop")
    print(profiler.dumps())
    profiler.dump()
    ev = json.load(open(args.file))["traceEvents"]
    cats = sorted(set(e.get("cat", "?") for e in ev))
    print("%d trace events written to %s (categories: %s)" % (len(ev), args.file, ", ".join(cats)))
    return ev


if __name__ == "__main__":
    main()
