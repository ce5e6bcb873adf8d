#!/usr/bin/env python
"""Model parallelism by placement: the two embedding tables of a matrix-factorisation model live on one device, the dense
tower on another; the executor inserts the cross-device copies (reference: example/model-parallel/matrix_factorization,
`mx.AttrScope(ctx_group=...)` + `group2ctxs`, src/executor/graph_executor.cc PlaceDevice).

    python examples/model-parallel/matrix_factorization.py [--gpus 0,1] [--epochs 3]

Synthetic ratings (a planted low-rank model + noise) stand in for MovieLens: there is no network on the build machines."""
import argparse
import logging
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mxb200 as mx  # noqa: E402


def build(factor, hidden, users, items):
    with mx.AttrScope(ctx_group="embed"):
        user, item = mx.sym.Variable("user"), mx.sym.Variable("item")
        u = mx.sym.Embedding(user, mx.sym.Variable("user_weight"), input_dim=users, output_dim=factor, name="user_embed")
        v = mx.sym.Embedding(item, mx.sym.Variable("item_weight"), input_dim=items, output_dim=factor, name="item_embed")
    with mx.AttrScope(ctx_group="dense"):
        u = mx.sym.FullyConnected(mx.sym.Activation(u, act_type="relu"), num_hidden=hidden, name="fc_user")
        v = mx.sym.FullyConnected(mx.sym.Activation(v, act_type="relu"), num_hidden=hidden, name="fc_item")
        pred = mx.sym.sum(u * v, axis=1)
        return mx.sym.LinearRegressionOutput(pred, mx.sym.Variable("score"), name="out")


def synthetic(users, items, n, rank, rng):
    pu, qi = rng.normal(0, 1, (users, rank)), rng.normal(0, 1, (items, rank))
    u, i = rng.integers(0, users, n), rng.integers(0, items, n)
    score = (pu[u] * qi[i]).sum(1) + rng.normal(0, 0.1, n)
    return u.astype(np.float32), i.astype(np.float32), score.astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="")
    ap.add_argument("--epochs", type=int, default=8)
    ap.add_argument("--batch-size", type=int, default=256)
    ap.add_argument("--users", type=int, default=200)
    ap.add_argument("--items", type=int, default=100)
    ap.add_argument("--samples", type=int, default=40000)
    ap.add_argument("--factor", type=int, default=32)
    ap.add_argument("--hidden", type=int, default=32)
    args = ap.parse_args()
    logging.basicConfig(level=logging.INFO)
    devs = [mx.gpu(int(g)) for g in args.gpus.split(",")] if args.gpus else [mx.cpu(0), mx.cpu(1)]
    group2ctx = {"embed": devs[0], "dense": devs[-1]}
    rng = np.random.default_rng(0)
    u, i, s = synthetic(args.users, args.items, args.samples, 8, rng)
    split = int(0.9 * args.samples)
    train = mx.io.NDArrayIter({"user": u[:split], "item": i[:split]}, {"score": s[:split]}, args.batch_size, shuffle=True)
    val = mx.io.NDArrayIter({"user": u[split:], "item": i[split:]}, {"score": s[split:]}, args.batch_size)
    net = build(args.factor, args.hidden, args.users, args.items)
    mod = mx.mod.Module(net, data_names=["user", "item"], label_names=["score"], context=[devs[0]], group2ctxs=[group2ctx])
    mod.fit(train, eval_data=val, eval_metric="rmse", optimizer="adam", optimizer_params={"learning_rate": 0.02},
            initializer=mx.init.Xavier(), num_epoch=args.epochs, batch_end_callback=mx.callback.Speedometer(args.batch_size, 50))
    rmse = mod.score(val, "rmse")[0][1]
    print("validation rmse %.4f (label std %.3f); embeddings on %s, dense tower on %s" % (rmse, float(s.std()), devs[0], devs[-1]))
    return rmse


if __name__ == "__main__":
    main()
