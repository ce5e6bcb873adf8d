#!/usr/bin/env python
"""Headline benchmark: ResNet-50 (v1b) ImageNet-shape training throughput, images/sec, whole job.

    python bench.py --gpus N --steps K --warmup W            # our framework (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference ...                     # the unmodified reference (unavailable here: see DESIGN.md)

Protocol (see task contract): W warm-up steps, then exactly K timed steps bracketed by barrier +
torch.cuda.synchronize(), timed with CUDA events on the launching stream, max over ranks.  Two timed regions:
  value : device-resident synthetic batch (the reference's own ``--benchmark 1`` SyntheticDataIter mode,
          example/image-classification/common/data.py:98-130)
  e2e   : every step copies that step's uint8 NHWC batch + labels host(pinned)->device and reads the loss back
          device->host, through the public gluon API.
Every step is a full training step: forward, loss, backward, fused allreduce + NAG-momentum update with fp32
master weights (nothing skipped).  Synthetic data, random-init weights (no network in this environment).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

BASELINE_V100_FP32_IMG_S = 363.69   # ResNet-50 train fp32 1xV100 batch 128 (docs/faq/perf.md:216)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=512,
                    help="per-GPU batch (weak scaling); 512 x 8 GPUs = the global batch 4096 of the reference recipe "
                         "(example/dynamic-training/run.sh:16)")
    ap.add_argument("--image", type=int, default=224)
    ap.add_argument("--classes", type=int, default=1000)
    ap.add_argument("--layers", type=int, default=50)
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def reference_arm(args):
    # The reference is an MXNet 1.3.1 source tree: python/setup.py needs a prebuilt libmxnet.so, whose native build
    # needs CUDA 9-era cuDNN APIs, OpenCV, BLAS and wget-ed ps-lite deps (Makefile:344, 3rdparty/ps-lite/make/deps.mk).
    # Both offline install attempts failed (baseline/ref_install*.log, DESIGN.md "Reference arm").
    print(json.dumps({"impl": "reference", "unavailable":
                      "reference is an MXNet-1.3.1 source tree; pip install fails: no setup.py at root and "
                      "python/setup.py cannot find libmxnet.so (native build needs cuDNN7/OpenCV/network)"}))
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import mxb200 as mx
    from mxb200 import ops as mxops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    B, S = args.batch, args.image
    from mxb200.gluon.model_zoo import vision
    from mxb200 import gluon, autograd
    ctx = mx.gpu(local_rank)
    net = vision.get_resnet("v1b", args.layers, classes=args.classes, ctx=ctx, dtype="bfloat16", layout="NHWC")
    net.hybridize()
    kv = mx.kv.create("device")
    # NAG momentum 0.9, wd 1e-4, multi-precision: the dynamic-training recipe (example/dynamic-training/run.sh:16)
    trainer = gluon.Trainer(net.collect_params(), "nag",
                            {"learning_rate": 0.1, "momentum": 0.9, "wd": 1e-4, "multi_precision": True},
                            kvstore=kv)
    loss_fn = gluon.loss.SoftmaxCrossEntropyLoss()
    global_batch = B * world

    g = torch.Generator().manual_seed(1234 + rank)
    host_img = torch.randint(0, 256, (B, S, S, 3), generator=g, dtype=torch.uint8).pin_memory()
    host_lab = torch.randint(0, args.classes, (B,), generator=g, dtype=torch.int64).pin_memory()
    mean = torch.tensor([123.68, 116.28, 103.53], device=dev).view(1, 1, 1, 3)
    istd = 1.0 / torch.tensor([58.395, 57.12, 57.375], device=dev).view(1, 1, 1, 3)

    def to_device_batch():
        img = host_img.to(dev, non_blocking=True)
        lab = host_lab.to(dev, non_blocking=True)
        x = ((img.float() - mean) * istd).to(torch.bfloat16)
        return mx.nd.from_torch(x, ctx), mx.nd.from_torch(lab, ctx)

    def train_step(x, y):
        with autograd.record():
            out = net(x)
            loss = loss_fn(out, y)
        loss.backward()
        trainer.step(global_batch)
        return loss

    res_x, res_y = to_device_batch()
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, after=None):
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        l0 = mxops.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if after is not None:
            after()                                # e.g. consume the last outstanding device->host read
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop()
        launches = mxops.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, clocks, launches

    def step_resident():
        train_step(res_x, res_y)

    # ---- end-to-end arm: the public input pipeline (mx.io.DevicePrefetcher: pinned uint8 batch -> H2D on a copy stream
    # -> normalise+cast kernel, one batch ahead of the compute) and a device->host read of every step's loss through a
    # pinned ring (the read of step i is issued in step i and consumed while step i+1 is in flight, so the host never
    # stalls the launch queue; every read completes inside the timed region)
    def source():
        while True:
            yield host_img, host_lab
    prefetch = iter(mx.io.DevicePrefetcher(source, ctx, mean=(123.68, 116.28, 103.53), std=(58.395, 57.12, 57.375),
                                           dtype="bfloat16"))
    ring = [(torch.zeros(1, dtype=torch.float32).pin_memory(), torch.cuda.Event()) for _ in range(2)]
    pending = []
    seen = [0.0, 0]

    def drain(keep):
        while len(pending) > keep:
            buf, ev = pending.pop(0)
            ev.synchronize()
            seen[0] = float(buf[0])
            seen[1] += 1

    def step_e2e():
        x, y = next(prefetch)
        loss = train_step(x, y)
        drain(1)                                   # at most one read outstanding: its ring slot is free again
        buf, ev = ring[(seen[1] + len(pending)) % 2]
        buf.copy_(loss.as_torch().detach().float().mean().reshape(1), non_blocking=True)
        ev.record()
        pending.append((buf, ev))

    for _ in range(max(3, args.warmup)):
        step_resident()
    ms, clocks, launches = timed(step_resident, args.steps)
    value = global_batch * args.steps / (ms / 1e3)

    e2e = None
    if not args.no_e2e:
        for _ in range(3):
            step_e2e()
        drain(0)
        n_before = seen[1]

        ms2, _c2, _l2 = timed(step_e2e, args.steps, after=lambda: drain(0))
        assert seen[1] - n_before == args.steps, "every timed step must have delivered its loss to the host"
        e2e = {"value": global_batch * args.steps / (ms2 / 1e3), "unit": "images/sec",
               "pipeline": "mx.io.DevicePrefetcher (pinned uint8 -> H2D on copy stream -> normalise kernel) + per-step async loss readback",
               "h2d_bytes_per_step": int(host_img.numel() * host_img.element_size() + host_lab.numel() * 8) * world,
               "d2h_bytes_per_step": 4 * world, "ms_per_step": ms2 / args.steps}
    trainer.check_health()
    if rank == 0:
        out = {
            "metric": "ResNet-50 training throughput", "value": value, "unit": "images/sec", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": value / BASELINE_V100_FP32_IMG_S,
            "dtype": "bf16", "data": "synthetic (random uint8 NHWC images, random labels; random-init weights)",
            "impl": "ours",
            "config": {"model": "resnet%d_v1b" % args.layers, "global_batch": global_batch, "per_gpu_batch": B,
                       "image": [S, S, 3], "classes": args.classes, "parallelism": "dp%d" % world,
                       "optimizer": "NAG mom=0.9 wd=1e-4, fp32 master weights (multi_precision), fused allreduce+update",
                       "kvstore": "device (fused P2P allreduce+update kernel)",
                       "l2_policy": "working set per step (GBs of activations) >> 126 MB L2; inputs not re-used from L2",
                       "baseline_ref": "BASELINE.md ResNet-50 train fp32 1xV100 batch128 = 363.69 img/s"},
            "clocks": clocks, "gpu_launches": launches, "final_loss": seen[0] if e2e else None,
        }
        if e2e:
            out["e2e"] = e2e
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
