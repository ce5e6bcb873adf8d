#!/usr/bin/env python
"""Elastic single-box training through the PUBLIC Gluon API: ``mx.kv.create('elastic_device')`` + ``gluon.Trainer``
(BASELINE.json config 3: ResNet-50, live add / remove of workers, dynamic-mini-batch learning-rate rule).

Worker tag local<k> -> GPU k; rank 0 edits the host file according to PLAN and the membership changes at the next epoch
boundary (the reference's protocol: example/dynamic-training/README.md, tools/launch.py host-file manager).  The per-GPU
batch is constant, so the global batch follows the membership and ``DynamicMiniBatchScheduler`` (linear rule + momentum
correction) is re-applied by ``KVStoreElasticDevice._membership_change_barrier`` -> ``Trainer._on_membership_change``.
The step is ``mx.gluon.TrainStep`` (CUDA-graph replay of forward + backward + fused P2P allreduce/update); graphs are
re-captured after every membership change because the arenas are re-mapped.

  python tools/launch.py -n 4 -s 1 --launcher local --elastic-training-enabled True \\
         python tools/elastic_gluon_worker.py <outdir> --plan 4,8,4 --layers 50 --image 224 --batch-per-gpu 512

Every worker writes <outdir>/out_<tag>.json: per epoch world / rank / images-per-second of the whole job (device-timed,
max over ranks) / learning rate / loss / weight checksum equality, and rank 0's re-init latency record per boundary.
"""
import argparse
import faulthandler
import json
import os
import struct
import sys

import torch

faulthandler.dump_traceback_later(900, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mxb200 as mx  # noqa: E402
from mxb200 import gluon  # noqa: E402
from mxb200.gluon.model_zoo import vision  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("outdir")
ap.add_argument("legacy", nargs="*")                       # round-1 positional form: plan steps global_batch
ap.add_argument("--plan", default="2,4,4,2,2")
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--layers", type=int, default=18)
ap.add_argument("--image", type=int, default=32)
ap.add_argument("--classes", type=int, default=10)
ap.add_argument("--batch-per-gpu", type=int, default=0, help="constant per-GPU batch (global batch follows the membership)")
ap.add_argument("--global-batch", type=int, default=64, help="used when --batch-per-gpu is 0: constant global batch")
ap.add_argument("--lr", type=float, default=0.02)
ap.add_argument("--no-graph", action="store_true")
args = ap.parse_args()
if args.legacy:
    args.plan = args.legacy[0]
    if len(args.legacy) > 1:
        args.steps = int(args.legacy[1])
    if len(args.legacy) > 2:
        args.global_batch = int(args.legacy[2])
outdir = args.outdir
plan = [int(v) for v in args.plan.split(",")]
tag = os.environ.get("MXB200_WORKER_TAG", "local0")
gpu = int(tag.replace("local", "") or 0)
is_new = os.environ.get("NEW_WORKER", "0") == "1"
begin = int(os.environ.get("EPOCH_BEGIN", "0")) if is_new else 0
host_file = os.environ.get("WORKER_HOST_FILE") or os.path.join(os.getcwd(), "host_worker")
ctx = mx.gpu(gpu)
torch.cuda.set_device(gpu)
dev = ctx.torch_device

kv = mx.kv.create("elastic_device")
thumb = args.image < 64
net = vision.get_resnet("v1b", args.layers, classes=args.classes, thumbnail=thumb, ctx=ctx, dtype="bfloat16", layout="NHWC",
                        seed=7 if not is_new else 99)
net.hybridize()
per_gpu = args.batch_per_gpu
base_world = plan[0]
sched = None
opt_params = {"learning_rate": args.lr, "momentum": 0.9, "wd": 1e-4, "multi_precision": True}
if per_gpu:
    # constant per-worker batch: linear lr scaling with the global batch + momentum correction at the boundary
    inner = mx.lr_scheduler.FactorScheduler(step=10 ** 9, factor=1.0)
    sched = mx.lr_scheduler.DynamicMiniBatchScheduler(inner, base_batch=per_gpu * base_world, mode="linear",
                                                      per_worker_batch=per_gpu)
    sched.base_lr = args.lr
    opt_params["lr_scheduler"] = sched
trainer = gluon.Trainer(net.collect_params(), "nag", opt_params, kvstore=kv)
trainer.init_kvstore()            # a newcomer adopts the live weights / momentum before its first forward pass
if sched is not None and is_new:
    sched.on_membership_change(per_gpu * kv.num_workers)     # a newcomer starts at the job's current global batch
L = gluon.loss.SoftmaxCrossEntropyLoss()
step = gluon.TrainStep(net, L, trainer, capture=not args.no_graph, warmup=1)
log = {"tag": tag, "gpu": gpu, "new": is_new, "begin": begin, "epochs": [], "config": vars(args)}
g = torch.Generator().manual_seed(123)
NS = 4 * max(per_gpu or args.global_batch, 64)
X = torch.randn(NS, args.image, args.image, 3, generator=g).bfloat16().to(dev)
Y = torch.randint(0, args.classes, (NS,), generator=g).to(dev)


def set_hosts(n):
    with open(host_file + ".tmp", "w") as f:
        f.write("\n".join("local%d" % i for i in range(n)) + "\n")
    os.rename(host_file + ".tmp", host_file)


def gather_max(v):
    vals = [struct.unpack("<d", b)[0] for b in kv.elastic_group.allgather(struct.pack("<d", float(v)))]
    return max(vals), vals


for epoch in range(begin, len(plan)):
    if not (is_new and epoch == begin):
        kv._membership_change_barrier({"EPOCH_BEGIN": str(epoch)})
        step.reset()
    world, rank = kv.num_workers, kv.rank
    global_batch = per_gpu * world if per_gpu else args.global_batch
    per = per_gpu or (global_batch // world)
    losses = []
    ev = []
    for it in range(args.steps):
        lo = ((epoch * args.steps + it) * per + rank * 7) % (NS - per)
        xb = mx.nd.from_torch(X[lo: lo + per], ctx)
        yb = mx.nd.from_torch(Y[lo: lo + per], ctx)
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        ev.append(e)
        loss = step(xb, yb, batch_size=global_batch)
        losses.append(loss.as_torch().detach().float().mean())
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    ev.append(e)
    torch.cuda.synchronize(dev)
    trainer.check_health()
    skip = min(3, args.steps - 1)                    # eager warm-up + graph capture steps are not steady state
    ms = ev[skip].elapsed_time(ev[-1]) / max(1, args.steps - skip)
    ms_max, _all = gather_max(ms)
    losses = [float(v) for v in losses]
    # trainable parameters only: BatchNorm running statistics are per-worker state between membership changes
    csum = float(sum(p.data().as_torch().detach().float().abs().sum() for p in net.collect_params().values()
                     if p.grad_req != "null"))
    sums = [struct.unpack("<d", b)[0] for b in kv.elastic_group.allgather(struct.pack("<d", csum))]
    rec = {"epoch": epoch, "rank": rank, "world": world, "global_batch": global_batch, "loss_first": losses[0],
           "loss_last": losses[-1], "ms_per_step_max_over_ranks": ms_max, "images_per_sec_job": global_batch / (ms_max / 1e3),
           "lr": trainer.learning_rate, "weights_identical": max(sums) - min(sums) <= 1e-6 * abs(sums[0]), "checksum": csum,
           "graph_replays": step.replays}
    if kv.reinit_log:
        rec["reinit"] = kv.reinit_log[-1]
        rec["lr_change"] = getattr(trainer, "last_lr_change", None)
    log["epochs"].append(rec)
    with open(os.path.join(outdir, "out_%s.json" % tag), "w") as f:
        json.dump(log, f)
    if rank == 0 and epoch + 1 < len(plan) and plan[epoch + 1] != world:
        set_hosts(plan[epoch + 1])
kv.close()
print("worker", tag, "done", flush=True)
