#!/usr/bin/env python
"""Sparse linear classification: CSR mini-batches, a row_sparse weight that lives in the kvstore, and per-batch
`row_sparse_pull` of only the rows the batch touches (reference: example/sparse/linear_classification/train.py, which
trains on the Avazu click log with `dist_sync` / `dist_async`; row ids -> `kvstore_utils` Unique -> sparse push / pull).

    python examples/sparse/linear_classification.py [--kvstore local|device|dist_sync] [--epochs 3]
    python tools/launch.py -n 2 -s 1 --launcher local python examples/sparse/linear_classification.py --kvstore dist_sync

Synthetic click log: 100 k features, ~30 active per example, labels from a planted sparse linear model."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mxb200 as mx  # noqa: E402
from mxb200 import nd  # noqa: E402


def synthetic(n, features, active, rng):
    idx = rng.integers(0, features, (n, active))
    val = np.ones((n, active), np.float32)
    w_true = np.zeros(features, np.float32)
    hot = rng.choice(features, features // 20, replace=False)
    w_true[hot] = rng.normal(0, 1.5, len(hot))
    logit = w_true[idx].sum(1)
    y = (logit + rng.normal(0, 0.3, n) > 0).astype(np.float32)
    indptr = np.arange(0, n * active + 1, active, dtype=np.int64)
    return idx.reshape(-1).astype(np.int64), val.reshape(-1), indptr, y


def batches(idx, val, indptr, y, bs, features, part, parts):
    n = len(y) // parts
    lo = part * n
    for b in range(lo, lo + n - bs + 1, bs):
        s, e = indptr[b], indptr[b + bs]
        csr = nd.sparse.csr_matrix((val[s:e], idx[s:e], indptr[b:b + bs + 1] - s), shape=(bs, features))
        yield csr, nd.array(y[b:b + bs]), np.unique(idx[s:e])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kvstore", default="local")
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--batch-size", type=int, default=512)
    ap.add_argument("--features", type=int, default=100000)
    ap.add_argument("--samples", type=int, default=40000)
    ap.add_argument("--lr", type=float, default=20.0)
    args = ap.parse_args()
    kv = mx.kv.create(args.kvstore)
    rng = np.random.default_rng(0)
    idx, val, indptr, y = synthetic(args.samples, args.features, 30, rng)
    # the weight is stored row_sparse; the optimizer runs where the store lives (servers for dist_*)
    weight = nd.sparse.zeros("row_sparse", (args.features, 1))
    kv.init("weight", weight)
    kv.set_optimizer(mx.optimizer.create("sgd", learning_rate=args.lr, rescale_grad=1.0 / args.batch_size))
    w_rows = nd.sparse.zeros("row_sparse", (args.features, 1))
    for epoch in range(args.epochs):
        tic, seen, loss_sum, correct = time.time(), 0, 0.0, 0
        for csr, label, rows in batches(idx, val, indptr, y, args.batch_size, args.features, kv.rank, kv.num_workers):
            row_ids = nd.array(rows, dtype="int64")
            kv.row_sparse_pull("weight", out=w_rows, row_ids=row_ids)          # only the rows this batch touches
            z = nd.sparse.dot(csr, w_rows).reshape((-1,))
            p = 1.0 / (1.0 + nd.exp(-z))
            err = p - label                                                  # d(logistic loss)/dz
            grad_dense_rows = nd.sparse.dot(csr, err.reshape((-1, 1)), transpose_a=True)   # [features, 1], non-zero on `rows`
            grad = nd.sparse.retain(grad_dense_rows.tostype("row_sparse"), row_ids)
            kv.push("weight", grad)                                          # sparse push; the store applies SGD
            pn, ln = p.asnumpy(), label.asnumpy()
            loss_sum += float(-(ln * np.log(pn + 1e-7) + (1 - ln) * np.log(1 - pn + 1e-7)).sum())
            correct += int(((pn > 0.5) == (ln > 0.5)).sum())
            seen += len(ln)
        print("worker %d epoch %d: %.0f samples/s, logloss %.4f, accuracy %.4f" % (kv.rank, epoch, seen / (time.time() - tic),
                                                                                  loss_sum / seen, correct / seen), flush=True)
    return correct / max(1, seen)


if __name__ == "__main__":
    main()
