#!/usr/bin/env python
"""Single-box elastic data-parallel training on GPUs (one process per GPU, worker set changes at epoch boundaries).

Launched through tools/launch.py --elastic-training-enabled True; worker tag ``local<k>`` runs on GPU k.  Rank 0 edits
the host file between barriers according to PLAN (epoch -> worker count), the scheduler starts / stops workers at the
next epoch boundary, and the fused P2P allreduce+update kernel keeps running over whatever GPUs are currently in the
job (IPC handles re-exchanged through the scheduler, ZeRO-1 optimizer state re-sharded, newcomers adopt live state).

Checks written to <outdir>/out_<tag>.json: world size per epoch, dense ranks, identical weights on all ranks after
every epoch (checksums all-gathered through the scheduler), momentum survives the re-shard, loss keeps decreasing.
"""
import faulthandler
import json
import os
import struct
import sys
import time

import torch

faulthandler.dump_traceback_later(240, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxb200.models import resnet as R  # noqa: E402
from mxb200.ops import functional as K  # noqa: E402
from mxb200.parallel.membership import ElasticDeviceTrainer, ElasticGroup  # noqa: E402

outdir = sys.argv[1]
plan = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "2,4,4,2,2").split(",")]   # workers wanted in epoch e
steps_per_epoch = int(sys.argv[3]) if len(sys.argv) > 3 else 6
global_batch = int(sys.argv[4]) if len(sys.argv) > 4 else 64
tag = os.environ.get("MXB200_WORKER_TAG", "local0")
gpu = int(tag.replace("local", "") or 0)
is_new = os.environ.get("NEW_WORKER", "0") == "1"
begin = int(os.environ.get("EPOCH_BEGIN", "0")) if is_new else 0
host_file = os.environ.get("WORKER_HOST_FILE") or os.path.join(os.getcwd(), "host_worker")
dev = torch.device("cuda", gpu)
torch.cuda.set_device(dev)

eg = ElasticGroup(tag=tag)
spec = R.ResNetSpec(18, classes=10, version="v1b", thumbnail=True)
P, A = R.init_resnet_params(spec, dtype=torch.bfloat16, device=dev, seed=0)
if is_new:      # a newcomer must not depend on its own initialisation: scramble it, live state comes from rank 0
    for v in P.values():
        v.add_(1.0)
aux = list(A.values())
sess = ElasticDeviceTrainer(eg, list(P.items()), dev, aux=aux)
log = {"tag": tag, "gpu": gpu, "new": is_new, "begin": begin, "epochs": []}

# fixed synthetic data set: 512 images, sharded by rank every epoch (global batch constant => per-rank batch changes)
g = torch.Generator().manual_seed(123)
X = torch.randn(512, 32, 32, 3, generator=g).bfloat16()
Y = torch.randint(0, 10, (512,), generator=g)


def set_hosts(n):
    lines = ["local%d" % i for i in range(n)]
    with open(host_file + ".tmp", "w") as f:
        f.write("\n".join(lines) + "\n")
    os.rename(host_file + ".tmp", host_file)


for epoch in range(begin, len(plan)):
    if not (is_new and epoch == begin):
        sess.epoch_boundary(epoch)
    W = sess.params()
    world, rank = sess.world, sess.rank
    per = global_batch // world
    losses = []
    for it in range(steps_per_epoch):
        lo = ((epoch * steps_per_epoch + it) * global_batch) % 512
        xb = X[lo + rank * per: lo + (rank + 1) * per].to(dev)
        yb = Y[lo + rank * per: lo + (rank + 1) * per].to(dev)
        logits = R.resnet_forward(spec, W, A, xb, training=True, G=sess.grad_sinks)
        loss = K.softmax_cross_entropy(logits, yb).sum() / per
        loss.backward()
        sess.step("nag", lr=0.02, wd=1e-4, momentum=0.9, rescale=1.0 / world)
        losses.append(float(loss.detach()))
    sess.tr.check()
    torch.cuda.synchronize(dev)
    csum = float(sum(v.detach().float().abs().sum() for v in W.values()))
    mom = sess.tr.arenas[0].state.get("mom")
    mom_norm = float(mom.abs().sum()) if mom is not None else 0.0      # the state tensor IS the owned slice
    sums = [struct.unpack("<d", b)[0] for b in eg.allgather(struct.pack("<d", csum))]
    log["epochs"].append({"epoch": epoch, "rank": rank, "world": world, "loss_first": losses[0], "loss_last": losses[-1],
                          "weights_identical": max(sums) - min(sums) <= 1e-6 * abs(sums[0]), "checksum": csum,
                          "own_momentum_abs_sum": mom_norm})
    with open(os.path.join(outdir, "out_%s.json" % tag), "w") as f:
        json.dump(log, f)
    if rank == 0 and epoch + 1 < len(plan) and plan[epoch + 1] != world:
        set_hosts(plan[epoch + 1])          # takes effect at the next epoch boundary
sess.close()
eg.close()
print("worker", tag, "done", flush=True)
