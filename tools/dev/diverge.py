"""2-GPU debug: which parameters differ across ranks after a few fused steps?"""
import os, sys, json
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", lrank); torch.cuda.set_device(dev)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
import mxb200 as mx
from mxb200 import gluon, autograd
which = sys.argv[1]
torch.manual_seed(3); mx.random.seed(3)
ctx = mx.gpu(lrank)
if which == "dense":
    net = gluon.nn.HybridSequential()
    with net.name_scope():
        net.add(gluon.nn.Dense(256, activation="relu", in_units=512), gluon.nn.Dense(512, in_units=256), gluon.nn.Dense(16, in_units=512))
    net.initialize(mx.init.Xavier(), ctx=ctx)
    mk = lambda s: (torch.randn(32, 512, device=dev, generator=torch.Generator(device=dev).manual_seed(50*s+rank)),
                    torch.randint(0, 16, (32,), device=dev, generator=torch.Generator(device=dev).manual_seed(7*s+rank)))
else:
    from mxb200.gluon.model_zoo import vision
    net = vision.get_resnet("v1b", 18, classes=64, ctx=ctx, dtype="bfloat16", layout="NHWC", thumbnail=True)
    net.hybridize()
    mk = lambda s: (torch.randn(16, 32, 32, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(50*s+rank)).bfloat16(),
                    torch.randint(0, 64, (16,), device=dev, generator=torch.Generator(device=dev).manual_seed(7*s+rank)))
tr = gluon.Trainer(net.collect_params(), "nag", {"learning_rate": 0.05, "momentum": 0.9, "wd": 1e-4, "multi_precision": True}, kvstore=mx.kv.create("device"))
L = gluon.loss.SoftmaxCrossEntropyLoss()
res = {}
for s in range(1, 5):
    x, y = mk(s)
    with autograd.record():
        l = L(net(mx.nd.from_torch(x, ctx)), mx.nd.from_torch(y, ctx))
    l.backward()
    tr.step(32 * world)
    torch.cuda.synchronize()
    bad = {}
    for name, p in net.collect_params().items():
        if p.grad_req == "null":
            continue
        t = p.data().as_torch().detach().float().contiguous()
        o = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(o, t)
        d = max((o[0] - o[r]).abs().max().item() for r in range(1, world))
        if d > 0:
            bad[name] = (d, tuple(t.shape))
    res[s] = bad
tr.check_health()
if rank == 0:
    info = [(ar.n, str(ar.wdtype), ar.one_shot(), bool(ar.g_mc), ar.slice_range()) for ar in tr._fused.arenas]
    print("RESULT", which, os.environ.get("MXB200_NVLS", "-"), os.environ.get("MXB200_SEQ_DEV", "-"), info, {k: (len(v), list(v.items())[:4]) for k, v in res.items()}, flush=True)
dist.destroy_process_group()
