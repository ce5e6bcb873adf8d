"""Worker of tests/test_multigpu.py: run under torchrun, one rank per GPU, scenario chosen by argv[1]; prints 'RESULT <json>'."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def setup():
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", lrank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    return rank, world, dev


def worst(dev, v):
    t = torch.tensor([float(v)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def elem_violation(got, ref, rtol, atol):
    """max over elements of |got-ref| / (atol + rtol*|ref|): <= 1 means every element is inside its own tolerance"""
    return ((got.float() - ref.float()).abs() / (atol + rtol * ref.float().abs())).max().item()


def scenario_optimizers(rank, world, dev):
    """two-shot / one-shot x SGD / NAG / Adam / LAMB across real GPUs against a single-device fp32 oracle, with the NVLS
    (multimem) and the P2P data paths (MXB200_NVLS is read when the arenas are allocated)."""
    from mxb200.parallel.trainer import DeviceTrainer
    out = {}
    for opt in ("sgd", "nag", "adam", "lamb"):
        for nelem in (6000, 300000):            # 24 KB (one-shot) and 1.2 MB (two-shot) of fp32 gradient
            torch.manual_seed(7)
            shapes = [nelem // 2, nelem // 3, nelem - nelem // 2 - nelem // 3]
            named = [("w%d" % i, (torch.randn(s, device=dev) * 0.1).bfloat16()) for i, s in enumerate(shapes)]
            w32 = [t.float().clone() for _, t in named]
            m = [torch.zeros_like(t) for t in w32]
            v = [torch.zeros_like(t) for t in w32]
            tr = DeviceTrainer(named, dev, rank, world)
            lr, wd, mom, b1, b2, eps = 0.05, 1e-3, 0.9, 0.9, 0.999, 1e-6
            for step in range(1, 4):
                gsum = []
                for i, (n, _) in enumerate(named):
                    gs = []
                    for r in range(world):      # every rank can reconstruct all ranks' gradients
                        g = torch.Generator(device=dev).manual_seed(1000 * step + 10 * i + r)
                        gs.append(torch.randn(shapes[i], device=dev, generator=g) * 0.01)
                    tr.grad_sinks[n].copy_(gs[rank])
                    gsum.append(torch.stack(gs).sum(0))
                kw = dict(lr=lr, wd=wd, rescale=1.0 / world)
                if opt in ("sgd", "nag"):
                    tr.step(opt, momentum=mom, **kw)
                else:
                    tr.step(opt, beta1=b1, beta2=b2, eps=eps, **kw)
                tr.check()
                for i in range(len(named)):
                    g = gsum[i] / world
                    if opt == "sgd":
                        m[i] = mom * m[i] - lr * wd * w32[i] - lr * g
                        w32[i] = w32[i] + m[i]
                    elif opt == "nag":
                        g = g + wd * w32[i]
                        m[i] = mom * m[i] + g
                        w32[i] = w32[i] - lr * (g + mom * m[i])
                    elif opt == "adam":
                        g = g + wd * w32[i]
                        m[i] = b1 * m[i] + (1 - b1) * g
                        v[i] = b2 * v[i] + (1 - b2) * g * g
                        lr_t = lr * (1 - b2 ** step) ** 0.5 / (1 - b1 ** step)
                        w32[i] = w32[i] - lr_t * m[i] / (v[i].sqrt() + eps)
                    else:
                        m[i] = b1 * m[i] + (1 - b1) * g
                        v[i] = b2 * v[i] + (1 - b2) * g * g
                        r_ = (m[i] / (1 - b1 ** step)) / ((v[i] / (1 - b2 ** step)).sqrt() + eps) + wd * w32[i]
                        wn, rn = w32[i].norm(), r_.norm()
                        ratio = (wn / rn) if (wn > 0 and rn > 0) else 1.0
                        w32[i] = w32[i] - lr * ratio * r_
            torch.cuda.synchronize()
            viol = 0.0
            for i, (n, _) in enumerate(named):
                got = tr.params()[n].detach()
                # bf16 weights: half an ulp of the stored value + a little accumulated-order noise
                viol = max(viol, elem_violation(got, w32[i], 2.0 ** -8, 2e-5))
                viol = max(viol, tr.grad_sinks[n].abs().max().item() * 1e9)      # gradients must have been consumed
            out["%s_%d" % (opt, nelem)] = worst(dev, viol)
            tr.close()
    return out


def scenario_checkpoint(rank, world, dev):
    """gluon.Trainer.save_states -> fresh process state -> load_states -> continue: identical to an uninterrupted run."""
    import tempfile
    import mxb200 as mx
    from mxb200 import gluon, autograd

    def build(seed):
        torch.manual_seed(seed)
        mx.random.seed(seed)
        ctx = mx.gpu(dev.index)
        net = gluon.nn.HybridSequential(prefix="ckpt_")
        with net.name_scope():
            net.add(gluon.nn.Dense(256, activation="relu", in_units=512), gluon.nn.Dense(512, in_units=256),
                    gluon.nn.Dense(16, in_units=512))
        net.initialize(mx.init.Xavier(), ctx=ctx)
        tr = gluon.Trainer(net.collect_params(), "nag", {"learning_rate": 0.05, "momentum": 0.9, "wd": 1e-4},
                           kvstore=mx.kv.create("device"))
        return ctx, net, tr

    def batch(step, ctx):
        g = torch.Generator(device=dev).manual_seed(50 * step + rank)
        x = torch.randn(32, 512, device=dev, generator=g)
        y = torch.randint(0, 16, (32,), device=dev, generator=g)
        return mx.nd.from_torch(x, ctx), mx.nd.from_torch(y, ctx)

    lossf = gluon.loss.SoftmaxCrossEntropyLoss()

    def run(net, tr, ctx, steps):
        for s in steps:
            x, y = batch(s, ctx)
            with autograd.record():
                l = lossf(net(x), y)
            l.backward()
            tr.step(32 * world)
        return [p.data().as_torch().detach().float().clone() for p in net.collect_params().values()]

    ctx, net, tr = build(3)
    run(net, tr, ctx, range(1, 4))
    d = tempfile.mkdtemp() if rank == 0 else None
    box = [d]
    dist.broadcast_object_list(box, src=0)
    d = box[0]
    fn = os.path.join(d, "trainer.states")
    wfn = os.path.join(d, "net.params")
    tr.save_states(fn + ".%d" % rank)            # collective; every rank writes its own (complete) copy
    if rank == 0:
        net.save_parameters(wfn)
    dist.barrier()
    ref = run(net, tr, ctx, range(4, 7))
    # "restart": new objects, same files (every rank loads rank 0's weights and ANOTHER rank's state file: all complete)
    ctx2, net2, tr2 = build(99)
    net2.load_parameters(wfn, ctx=ctx2)
    x, y = batch(0, ctx2)
    with autograd.record():
        l = lossf(net2(x), y)
    l.backward()
    tr2.step(32 * world)                          # fuses the trainer (lazy), then the state overwrites everything
    net2.load_parameters(wfn, ctx=ctx2)
    tr2.load_states(fn + ".%d" % ((rank + 1) % world))
    got = run(net2, tr2, ctx2, range(4, 7))
    diff = max((a - b).abs().max().item() for a, b in zip(ref, got))
    sharded = bool(tr._fused is not None and any(ar.slice_range() != (0, ar.n) for ar in tr._fused.arenas))
    state_elems = sum(st.numel() for ar in tr._fused.arenas for st in ar.state.values()) if tr._fused else 0
    full_elems = sum(ar.n * len(ar.state) for ar in tr._fused.arenas) if tr._fused else 0
    return {"resume_max_abs_diff": worst(dev, diff), "fused": tr._fused is not None, "sharded": sharded,
            "state_elems": state_elems, "full_state_elems": full_elems}


def scenario_graph(rank, world, dev):
    """mx.gluon.TrainStep (CUDA-graph replay incl. the cross-GPU fused update) keeps all ranks' weights identical and follows
    the eager trajectory."""
    import mxb200 as mx
    from mxb200 import gluon
    from mxb200.gluon.model_zoo import vision
    res = {}
    for capture in (False, True):
        torch.manual_seed(0)
        mx.random.seed(0)
        ctx = mx.gpu(dev.index)
        net = vision.get_resnet("v1b", 18, classes=64, ctx=ctx, dtype="bfloat16", layout="NHWC", thumbnail=True)
        net.hybridize()
        tr = gluon.Trainer(net.collect_params(), "nag", {"learning_rate": 0.002, "momentum": 0.9, "wd": 1e-4,
                                                          "multi_precision": True}, kvstore=mx.kv.create("device"))
        step = gluon.TrainStep(net, gluon.loss.SoftmaxCrossEntropyLoss(), tr, batch_size=16 * world, capture=capture, warmup=2)
        losses = []
        for s in range(8):
            g = torch.Generator(device=dev).manual_seed(100 * s + rank)
            x = torch.randn(16, 32, 32, 3, device=dev, generator=g).bfloat16()
            y = torch.randint(0, 64, (16,), device=dev, generator=g)
            l = step(mx.nd.from_torch(x, ctx), mx.nd.from_torch(y, ctx))
            losses.append(float(l.as_torch().float().mean().item()))
        tr.check_health()
        # trainable parameters only: BatchNorm running statistics are per-rank state
        csum = sum(p.data().as_torch().float().abs().sum().item() for p in net.collect_params().values() if p.grad_req != "null")
        t = torch.tensor([csum], device=dev, dtype=torch.float64)
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        res["graph" if capture else "eager"] = {"losses": losses, "replays": step.replays,
                                                 "weights_identical": bool((hi - lo).item() <= 1e-9 * abs(hi.item()))}
    res["max_loss_diff"] = max(abs(a - b) for a, b in zip(res["eager"]["losses"], res["graph"]["losses"]))
    return res


def main():
    rank, world, dev = setup()
    name = sys.argv[1]
    out = {"optimizers": scenario_optimizers, "checkpoint": scenario_checkpoint, "graph": scenario_graph}[name](rank, world, dev)
    from mxb200.parallel import fused
    out["nvls_env"] = os.environ.get("MXB200_NVLS", "default")
    if rank == 0:
        print("RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
