"""GPU check: TrainStep (CUDA-graph replay) follows the same loss trajectory as the eager step."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import mxb200 as mx
from mxb200 import gluon, autograd
from mxb200.gluon.model_zoo import vision


def run(capture, steps=8, opt="nag"):
    torch.manual_seed(0)
    mx.random.seed(0)
    ctx = mx.gpu(0)
    net = vision.get_resnet("v1b", 50, classes=100, ctx=ctx, dtype="bfloat16", layout="NHWC")
    net.hybridize()
    sched = mx.lr_scheduler.FactorScheduler(step=2, factor=0.5)
    kw = {"learning_rate": 0.002 if opt == "nag" else 0.0005, "wd": 1e-4, "multi_precision": True, "lr_scheduler": sched}
    if opt == "nag":
        kw["momentum"] = 0.9
    tr = gluon.Trainer(net.collect_params(), opt, kw, kvstore="device")
    lf = gluon.loss.SoftmaxCrossEntropyLoss()
    step = gluon.TrainStep(net, lf, tr, batch_size=16, capture=capture, warmup=2)
    g = torch.Generator(device="cuda").manual_seed(1)
    xs = [torch.randn(16, 64, 64, 3, device="cuda", generator=g).bfloat16() for _ in range(steps)]
    ys = [torch.randint(0, 100, (16,), device="cuda", generator=g) for _ in range(steps)]
    out = []
    for x, y in zip(xs, ys):
        l = step(mx.nd.from_torch(x, ctx), mx.nd.from_torch(y, ctx))
        out.append(float(l.as_torch().float().mean().item()))
    return out, step.replays


res = {}
for opt in ("nag", "adam"):
    a, _ = run(False, opt=opt)
    b, rep = run(True, opt=opt)
    res[opt] = {"eager": a, "graph": b, "replays": rep, "max_abs_diff": max(abs(x - y) for x, y in zip(a, b))}
print(json.dumps(res))
ok = all(v["max_abs_diff"] < 0.15 and v["replays"] >= 5 for v in res.values())
print("GRAPH_CHECK", "OK" if ok else "FAIL")
sys.exit(0 if ok else 1)
