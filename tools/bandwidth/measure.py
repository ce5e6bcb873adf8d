#!/usr/bin/env python
"""KVStore push+pull bandwidth (the reference's tools/bandwidth/measure.py, same CLI and the same formula:
GB/s per GPU = total_MB * 2*(n-1)/n / seconds / 1e3, tools/bandwidth/measure.py:138) on the B200 framework.

Workload: every parameter tensor of the named model-zoo network (fp32 gradients), one push + one pull per tensor
and per iteration through ``mx.kv`` (``--mode kv``, the reference's path), or the whole set in ONE fused
allreduce(+update) kernel over the symmetric arenas (``--mode fused``, what gluon.Trainer / Module use here).

Single process (``--gpus 0,1``: one process drives several devices, like the reference) or one process per GPU under
torchrun (WORLD_SIZE>1; then --gpus is ignored and each rank uses LOCAL_RANK).  Device-timed with CUDA events
(max over ranks) in addition to the reference's host timer.
"""
import argparse
import json
import logging
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mxb200 as mx  # noqa: E402

logger = logging.getLogger()
logger.setLevel(logging.INFO)


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description="command for benchmark kv-store performance")
    parser.add_argument('--network', type=str, default="resnet", help='the neural network to test')
    parser.add_argument('--gpus', type=str, default='0,1', help='the gpus to be used, e.g "0,1,2,3"')
    parser.add_argument('--num-layers', type=int, default=152, help='number of layers, can be used for resnet')
    parser.add_argument('--kv-store', type=str, default='device', help='the kvstore type')
    parser.add_argument('--num-batches', type=int, default=5, help='number of batches to run')
    parser.add_argument('--disp-batches', type=int, default=1, help='show averaged results for every n batches')
    parser.add_argument('--test-results', type=int, default=1, help='if or not evalute the results correctness')
    parser.add_argument('--image-shape', type=str, default='3,224,224', help='input data shape')
    parser.add_argument('--num-classes', type=int, default=1000, help='number of classes')
    parser.add_argument('--optimizer', type=str, default='None', help='the optimizer set to kvstore. None means no optimizer')
    parser.add_argument('--gc-type', type=str, default='none', help='type of gradient compression')
    parser.add_argument('--mode', type=str, default='kv', choices=['kv', 'fused'], help='per-key push/pull or fused arena step')
    parser.add_argument('--cpu', action='store_true', help='run on CPU contexts (functional check without GPUs)')
    parser.add_argument('--json', type=str, default='', help='write the summary to this file')
    args = parser.parse_args(argv)
    logging.info(args)
    return args


def get_shapes(network, num_layers, num_classes):
    """(name, shape) of every trainable parameter of the model-zoo network (no bind needed)."""
    from mxb200.gluon.model_zoo import vision
    name = network
    if network == "resnet":
        name = "resnet%d_v1" % num_layers
    elif network in ("inception-v3", "inception_v3", "inceptionv3"):
        name = "inceptionv3"
    net = vision.get_model(name, classes=num_classes)
    net.initialize(ctx=mx.cpu())
    x = mx.nd.zeros((1, 3, 299, 299) if "inception" in name else (1, 3, 224, 224))
    with mx.autograd.predict_mode():
        net(x)
    return [(n, tuple(p.shape)) for n, p in net.collect_params().items() if p.grad_req != "null"]


class Results(object):
    def __init__(self, iter=0, time=0.0, bandwidth=0.0, error=0.0, device_ms=0.0):
        self.iter, self.time, self.bandwidth, self.error, self.device_ms = iter, time, bandwidth, error, device_ms


def error(gpu_res, cpu_res):
    return sum(float(np.abs(g - c).sum()) for g, c in zip(gpu_res, cpu_res)) / max(1e-30, sum(float(np.abs(c).sum()) for c in cpu_res))


def run(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lrank = int(os.environ.get("LOCAL_RANK", "0"))
    use_gpu = torch.cuda.is_available() and not args.cpu
    if world > 1:
        import torch.distributed as dist
        if use_gpu:
            torch.cuda.set_device(lrank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
        else:
            dist.init_process_group("gloo")
        devs = [mx.gpu(lrank) if use_gpu else mx.cpu()]
        n_total = world
    else:
        ids = [int(i) for i in args.gpus.split(',')]
        devs = [mx.gpu(i) for i in ids] if use_gpu else [mx.cpu(i) for i in ids]
        n_total = len(devs)
    shapes = get_shapes(args.network, args.num_layers, args.num_classes)
    size_mb = float(sum(int(np.prod(s)) for _, s in shapes) * 4) / 1e6
    logging.info('num of arrays = %d, total size = %f MB', len(shapes), size_mb)
    res = []

    def sync():
        if use_gpu:
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([v], device="cuda" if use_gpu else "cpu", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    if args.mode == 'fused':
        from mxb200.parallel.trainer import DeviceTrainer
        assert world > 1 or len(devs) == 1, "--mode fused runs one process per GPU (torchrun)"
        dev = torch.device("cuda", lrank) if use_gpu else torch.device("cpu")
        named = [(n, torch.zeros(s, dtype=torch.float32, device=dev)) for n, s in shapes]
        tr = DeviceTrainer(named, dev, rank, world)
        for b in range(args.num_batches + 1):
            for n, _ in named:
                tr.grad_sinks[n].fill_(float(rank + 1))
            sync()
            tic = time.time()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            tr.step("sgd", lr=1.0, wd=0.0, momentum=0.0, rescale=1.0)   # w -= sum of gradients
            e1.record()
            sync()
            toc = max_over_ranks(time.time() - tic)
            dms = max_over_ranks(e0.elapsed_time(e1))
            if b == 0:
                continue
            err = 0.0
            if args.test_results:
                expect = -float(b + 1) * world * (world + 1) / 2.0      # the warm-up iteration was applied as well
                err = max(abs(float(tr.params()[n].flatten()[0]) - expect) / abs(expect) for n, _ in named[:4])
            res.append(Results(b, toc, size_mb * 2 * (n_total - 1) / max(n_total, 1) / (dms / 1e3) / 1e3, err, dms))
            if rank == 0:
                logging.info('iter %d, %f sec (device %f ms), %f GB/sec per gpu, error %f', b, toc, dms, res[-1].bandwidth, err)
        tr.close()
    else:
        kv = mx.kv.create(args.kv_store)
        if args.gc_type != 'none':
            kv.set_gradient_compression({'type': args.gc_type})
        if args.optimizer != 'None':
            kv.set_optimizer(mx.optimizer.create(args.optimizer, learning_rate=0.0))
        grads = [[mx.nd.ones(s, ctx=d) * (i + 1 + rank) for i, d in enumerate(devs)] for _, s in shapes]
        weights = [[mx.nd.zeros(s, ctx=d) for d in devs] for _, s in shapes]
        for i, (_, s) in enumerate(shapes):
            kv.init(i, mx.nd.zeros(s))
        for b in range(args.num_batches + 1):
            sync()
            tic = time.time()
            if use_gpu:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            for i, g in enumerate(grads):
                kv.push(i, g, priority=-i)
            for i, w in enumerate(weights):
                kv.pull(i, w, priority=-i)
            for ws in weights:
                for w in ws:
                    w.wait_to_read()
            if use_gpu:
                e1.record()
            sync()
            toc = max_over_ranks(time.time() - tic)
            dms = max_over_ranks(e0.elapsed_time(e1)) if use_gpu else toc * 1e3
            if b == 0:
                continue
            err = 0.0
            if args.test_results and args.optimizer == 'None' and args.gc_type == 'none':
                if world > 1:
                    expect = sum(r + 1 for r in range(world))
                else:
                    expect = sum(i + 1 for i in range(len(devs)))
                err = max(abs(float(w[0].asnumpy().flat[0]) - expect) / expect for w in weights[:4])
            res.append(Results(b, toc, size_mb * 2 * (n_total - 1) / max(n_total, 1) / toc / 1e3, err, dms))
            if rank == 0:
                logging.info('iter %d, %f sec (device %f ms), %f GB/sec per gpu, error %f', b, toc, dms, res[-1].bandwidth, err)
    if rank == 0 and args.json:
        with open(args.json, "w") as f:
            json.dump({"network": args.network, "num_layers": args.num_layers, "arrays": len(shapes), "size_mb": size_mb,
                       "n": n_total, "mode": args.mode, "kv_store": args.kv_store,
                       "iters": [{"iter": r.iter, "sec": r.time, "device_ms": r.device_ms, "gbps_per_gpu": r.bandwidth,
                                  "error": r.error} for r in res]}, f, indent=1)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return res


if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO)
    run(parse_args())
