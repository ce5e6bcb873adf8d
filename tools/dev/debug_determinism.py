import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mxb200 as mx
from mxb200 import gluon, autograd
from mxb200.gluon.model_zoo import vision


def run(fuse_first):
    torch.manual_seed(0)
    mx.random.seed(0)
    ctx = mx.gpu(0)
    net = vision.get_resnet("v1b", 18, classes=10, thumbnail=True, ctx=ctx, dtype="bfloat16", layout="NHWC")
    net.hybridize()
    tr = gluon.Trainer(net.collect_params(), "sgd", {"learning_rate": 0.0, "multi_precision": True}, kvstore="device")
    lf = gluon.loss.SoftmaxCrossEntropyLoss()
    g = torch.Generator(device="cuda").manual_seed(3)
    x0 = mx.nd.from_torch(torch.randn(8, 32, 32, 3, device="cuda", generator=g).bfloat16(), ctx)
    y = mx.nd.from_torch(torch.randint(0, 10, (8,), device="cuda", generator=g), ctx)
    w0 = {n.split("_", 1)[1]: p.data().as_torch().detach().float().clone() for n, p in net.collect_params().items()}
    outs = []
    for it in range(3):
        with autograd.record():
            out = net(x0)
            l = lf(out, y)
        l.backward()
        grads = {n.split("_", 1)[1]: p.grad().as_torch().detach().float().clone() for n, p in net.collect_params().items() if p.grad_req != "null"}
        outs.append((out.as_torch().detach().float().clone(), grads))
        tr.step(8)       # lr = 0: weights unchanged, but the trainer fuses after the first step (arena sinks from then on)
    return w0, outs


def nrm(d):
    return sum(float((v.double() ** 2).sum()) for v in d.values()) ** 0.5


(wa, oa), (wb, ob) = run(True), run(True)
print("init weights differ:", nrm({k: wa[k] - wb[k] for k in wa}))
for it in range(3):
    fa, ga = oa[it]
    fb, gb = ob[it]
    print("iter %d: forward out diff %.4g (|out| %.4g)   grads: ||g|| %.4g  run-to-run diff %.4g   vs iter0 of same run %.4g" % (
        it, float((fa - fb).abs().max()), float(fa.abs().max()), nrm(ga), nrm({k: ga[k] - gb[k] for k in ga}),
        nrm({k: ga[k] - oa[0][1][k] for k in ga})))
    worst = sorted(((float((ga[k] - gb[k]).abs().max() / (ga[k].abs().max() + 1e-9)), k) for k in ga), reverse=True)[:4]
    print("   worst params:", ["%s %.3g" % (k, e) for e, k in worst])
