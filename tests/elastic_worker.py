"""Worker used by tests/test_elastic.py: a dist_sync job whose membership changes at epoch boundaries."""
import json
import os
import sys

import faulthandler

import numpy as np

faulthandler.dump_traceback_later(90, exit=True)   # a hung collective must fail the test, not stall it

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mxb200 as mx  # noqa: E402

outdir = sys.argv[1]
tag = os.environ.get("MXB200_WORKER_TAG", "?")
is_new = os.environ.get("NEW_WORKER", "0") == "1"
begin = int(os.environ.get("EPOCH_BEGIN", "0")) if is_new else 0
host_file = os.path.join(outdir, "host_worker")

kv = mx.kv.create("dist_sync")
shape = (4,)
kv.init("w", mx.nd.zeros(shape))
kv.init("bn_mean", mx.nd.zeros(shape), exclude_update=True)
if not is_new:
    kv.set_optimizer(mx.optimizer.create("test", rescale_grad=1.0))
out = mx.nd.zeros(shape)
if is_new:
    kv.pull("w", out=out)          # a newcomer obtains live state from the servers
log = {"tag": tag, "new": is_new, "begin": begin, "epochs": [], "pulled_on_join": out.asnumpy().tolist()}
expected_total = None
for epoch in range(begin, 5):
    if not (is_new and epoch == begin):
        kv._membership_change_barrier({"EPOCH_BEGIN": str(epoch)})
    n, rank = kv.num_workers, kv.rank
    kv.push("w", mx.nd.ones(shape))
    kv.pull("w", out=out)
    kv.push("bn_mean", mx.nd.ones(shape) * (rank + 1))
    aux = mx.nd.zeros(shape)
    kv.pull("bn_mean", out=aux)
    log["epochs"].append({"epoch": epoch, "rank": rank, "n": n, "w": float(out.asnumpy()[0]), "aux": float(aux.asnumpy()[0])})
    with open(os.path.join(outdir, "out_%s.json" % tag), "w") as f:
        json.dump(log, f)
    # membership edits are made by rank 0 between barriers (they take effect at the NEXT barrier)
    if rank == 0 and epoch == 0:
        with open(host_file, "a") as f:
            f.write("local2\n")
    if rank == 0 and epoch == 2:
        lines = [l for l in open(host_file).read().split() if l != "local2"]
        with open(host_file + ".tmp", "w") as f:
            f.write("\n".join(lines) + "\n")
        os.rename(host_file + ".tmp", host_file)
print("worker", tag, "done", flush=True)
