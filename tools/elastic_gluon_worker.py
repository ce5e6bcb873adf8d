#!/usr/bin/env python
"""Elastic single-box training through the PUBLIC Gluon API: ``mx.kv.create('elastic_device')`` + ``gluon.Trainer``.

Same scenario as tools/elastic_gpu_worker.py (worker tag local<k> -> GPU k, rank 0 edits the host file according to PLAN,
membership changes at epoch boundaries), but the model is a Gluon network and the step is ``trainer.step`` -- the fused
P2P allreduce+update kernel over whatever GPUs are currently in the job.  Launch:

  python tools/launch.py -n 2 -s 1 --launcher local --elastic-training-enabled True \\
         python tools/elastic_gluon_worker.py <outdir> 2,4,4,2,2
"""
import faulthandler
import json
import os
import struct
import sys

import torch

faulthandler.dump_traceback_later(300, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mxb200 as mx  # noqa: E402
from mxb200 import autograd, gluon  # noqa: E402
from mxb200.gluon.model_zoo import vision  # noqa: E402

outdir = sys.argv[1]
plan = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "2,4,4,2,2").split(",")]
steps_per_epoch = int(sys.argv[3]) if len(sys.argv) > 3 else 6
global_batch = int(sys.argv[4]) if len(sys.argv) > 4 else 64
tag = os.environ.get("MXB200_WORKER_TAG", "local0")
gpu = int(tag.replace("local", "") or 0)
is_new = os.environ.get("NEW_WORKER", "0") == "1"
begin = int(os.environ.get("EPOCH_BEGIN", "0")) if is_new else 0
host_file = os.environ.get("WORKER_HOST_FILE") or os.path.join(os.getcwd(), "host_worker")
ctx = mx.gpu(gpu)
torch.cuda.set_device(gpu)

kv = mx.kv.create("elastic_device")
net = vision.get_resnet("v1b", 18, classes=10, thumbnail=True, ctx=ctx, dtype="bfloat16", layout="NHWC", seed=7 if not is_new else 99)
net.hybridize()
trainer = gluon.Trainer(net.collect_params(), "nag", {"learning_rate": 0.02, "momentum": 0.9, "wd": 1e-4, "multi_precision": True},
                        kvstore=kv)
trainer.init_kvstore()            # a newcomer adopts the live weights / momentum before its first forward pass
L = gluon.loss.SoftmaxCrossEntropyLoss()
log = {"tag": tag, "gpu": gpu, "new": is_new, "begin": begin, "epochs": []}
g = torch.Generator().manual_seed(123)
X = torch.randn(512, 32, 32, 3, generator=g).bfloat16()
Y = torch.randint(0, 10, (512,), generator=g)


def set_hosts(n):
    with open(host_file + ".tmp", "w") as f:
        f.write("\n".join("local%d" % i for i in range(n)) + "\n")
    os.rename(host_file + ".tmp", host_file)


for epoch in range(begin, len(plan)):
    if not (is_new and epoch == begin):
        kv._membership_change_barrier({"EPOCH_BEGIN": str(epoch)})
    world, rank = kv.num_workers, kv.rank
    per = global_batch // world
    losses = []
    for it in range(steps_per_epoch):
        lo = ((epoch * steps_per_epoch + it) * global_batch) % 512
        xb = mx.nd.from_torch(X[lo + rank * per: lo + (rank + 1) * per].to(ctx.torch_device), ctx)
        yb = mx.nd.from_torch(Y[lo + rank * per: lo + (rank + 1) * per].to(ctx.torch_device), ctx)
        with autograd.record():
            loss = L(net(xb), yb)
        loss.backward()
        trainer.step(global_batch)
        losses.append(float(loss.as_torch().detach().float().mean()))
    trainer.check_health()
    # trainable parameters only: BatchNorm running statistics are per-worker state between membership changes
    csum = float(sum(p.data().as_torch().detach().float().abs().sum() for p in net.collect_params().values()
                     if p.grad_req != "null"))
    sums = [struct.unpack("<d", b)[0] for b in kv.elastic_group.allgather(struct.pack("<d", csum))]
    log["epochs"].append({"epoch": epoch, "rank": rank, "world": world, "loss_first": losses[0], "loss_last": losses[-1],
                          "weights_identical": max(sums) - min(sums) <= 1e-6 * abs(sums[0]), "checksum": csum})
    with open(os.path.join(outdir, "out_%s.json" % tag), "w") as f:
        json.dump(log, f)
    if rank == 0 and epoch + 1 < len(plan) and plan[epoch + 1] != world:
        set_hosts(plan[epoch + 1])
kv.close()
print("worker", tag, "done", flush=True)
