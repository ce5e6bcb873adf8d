#!/usr/bin/env python
"""Symbolic ImageNet training through ``Module.fit`` (reference: example/image-classification/train_imagenet.py +
common/fit.py, the script behind the scaling table quoted in BASELINE.md: ``--kv-store dist_device_sync``, fixed batch
per GPU).  ``--benchmark 1`` uses synthetic data like the reference (common/data.py:98-130)."""
import argparse
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mxb200 as mx  # noqa: E402
from mxb200.gluon.model_zoo import vision  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="train imagenet-1k")
    ap.add_argument("--network", default="resnet")
    ap.add_argument("--num-layers", type=int, default=50)
    ap.add_argument("--gpus", default="")
    ap.add_argument("--kv-store", default="device")
    ap.add_argument("--num-epochs", type=int, default=1)
    ap.add_argument("--lr", type=float, default=0.1)
    ap.add_argument("--lr-factor", type=float, default=0.1)
    ap.add_argument("--lr-step-epochs", default="30,60")
    ap.add_argument("--optimizer", default="sgd")
    ap.add_argument("--mom", type=float, default=0.9)
    ap.add_argument("--wd", type=float, default=1e-4)
    ap.add_argument("--batch-size", type=int, default=128, help="per process (split over --gpus)")
    ap.add_argument("--disp-batches", type=int, default=20)
    ap.add_argument("--model-prefix", default="")
    ap.add_argument("--load-epoch", type=int, default=None)
    ap.add_argument("--top-k", type=int, default=0)
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--benchmark", type=int, default=0)
    ap.add_argument("--data-train", default="")
    ap.add_argument("--data-val", default="")
    ap.add_argument("--image-shape", default="3,224,224")
    ap.add_argument("--num-classes", type=int, default=1000)
    ap.add_argument("--num-examples", type=int, default=1281167)
    ap.add_argument("--gc-type", default="none")
    ap.add_argument("--gc-threshold", type=float, default=0.5)
    args = ap.parse_args()
    logging.basicConfig(level=logging.INFO, format="%(asctime)-15s Node[0] %(message)s")
    logging.info("start with arguments %s", args)

    kv = mx.kv.create(args.kv_store)
    if args.gc_type != "none":
        kv.set_gradient_compression({"type": args.gc_type, "threshold": args.gc_threshold})
    if args.gpus and "LOCAL_RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        devs = [mx.gpu(int(os.environ["LOCAL_RANK"]))]      # torchrun: one process per GPU, --gpus only says "use GPUs"
    else:
        devs = [mx.gpu(int(i)) for i in args.gpus.split(",")] if args.gpus else [mx.cpu()]
    shape = tuple(int(v) for v in args.image_shape.split(","))
    name = ("resnet%d_v1" % args.num_layers) if args.network == "resnet" else args.network
    net = vision.get_model(name, classes=args.num_classes)
    data = mx.sym.var("data")
    if args.dtype != "float32":
        data = mx.sym.Cast(data=data, dtype=args.dtype)
        net.cast(args.dtype)
    out = net(data)
    if args.dtype != "float32":
        out = mx.sym.Cast(data=out, dtype="float32")
    sym = mx.sym.SoftmaxOutput(out, name="softmax")

    epoch_size = max(1, args.num_examples // args.batch_size // kv.num_workers)
    if args.benchmark or not args.data_train:
        train = mx.io.SyntheticDataIter(args.num_classes, (args.batch_size,) + shape, min(epoch_size, 50), dtype="float32")
        val = None
    else:
        train = mx.io.ImageRecordIter(path_imgrec=args.data_train, data_shape=shape, batch_size=args.batch_size, shuffle=True,
                                      rand_crop=True, rand_mirror=True, num_parts=kv.num_workers, part_index=kv.rank,
                                      ctx=devs[0] if (len(devs) == 1 and devs[0].device_type == "gpu") else None)
        val = mx.io.ImageRecordIter(path_imgrec=args.data_val, data_shape=shape, batch_size=args.batch_size,
                                    num_parts=kv.num_workers, part_index=kv.rank) if args.data_val else None
    steps = [epoch_size * int(e) for e in args.lr_step_epochs.split(",") if int(e) > 0]
    sched = mx.lr_scheduler.MultiFactorScheduler(step=steps, factor=args.lr_factor) if steps else None
    opt_params = {"learning_rate": args.lr, "wd": args.wd, "lr_scheduler": sched, "multi_precision": args.dtype != "float32"}
    if args.optimizer in ("sgd", "nag", "dcasgd", "signum", "lbsgd"):
        opt_params["momentum"] = args.mom
    arg_params = aux_params = None
    begin = 0
    if args.load_epoch is not None and args.model_prefix:
        _s, arg_params, aux_params = mx.model.load_checkpoint(args.model_prefix, args.load_epoch)
        begin = args.load_epoch
    metrics = ["accuracy"] + ([mx.metric.create("top_k_accuracy", top_k=args.top_k)] if args.top_k > 0 else [])
    mod = mx.mod.Module(sym, context=devs)
    mod.fit(train, eval_data=val, eval_metric=metrics, begin_epoch=begin, num_epoch=args.num_epochs, kvstore=kv,
            optimizer=args.optimizer, optimizer_params=opt_params, arg_params=arg_params, aux_params=aux_params,
            initializer=mx.init.Xavier(rnd_type="gaussian", factor_type="in", magnitude=2), allow_missing=True,
            batch_end_callback=mx.callback.Speedometer(args.batch_size, args.disp_batches),
            epoch_end_callback=mx.callback.do_checkpoint(args.model_prefix) if args.model_prefix else None)
