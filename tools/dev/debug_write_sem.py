import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mxb200 as mx
from mxb200 import gluon, autograd
from mxb200.gluon.model_zoo import vision


def run(skip_first):
    torch.manual_seed(0)
    mx.random.seed(0)
    ctx = mx.gpu(0)
    net = vision.get_resnet("v1b", 18, classes=10, thumbnail=True, ctx=ctx, dtype="bfloat16", layout="NHWC")
    net.hybridize()
    tr = gluon.Trainer(net.collect_params(), "sgd", {"learning_rate": 0.05, "multi_precision": True}, kvstore="device")
    lf = gluon.loss.SoftmaxCrossEntropyLoss()
    g = torch.Generator(device="cuda").manual_seed(3)
    x0 = mx.nd.from_torch(torch.randn(8, 32, 32, 3, device="cuda", generator=g).bfloat16(), ctx)
    x1 = mx.nd.from_torch(torch.randn(8, 32, 32, 3, device="cuda", generator=g).bfloat16(), ctx)
    y = mx.nd.from_torch(torch.randint(0, 10, (8,), device="cuda", generator=g), ctx)
    with autograd.record():
        l = lf(net(x0), y)
    l.backward()
    tr.step(8)
    gskip = None
    if skip_first:
        with autograd.record():
            l = lf(net(x0 * 3.0), y)
        l.backward()
        gskip = {}
        for ar in tr._fused.arenas:
            print("after skipped backward: dirty", ar.dirty, "bw_epoch", ar.bw_epoch, "write", ar.write_semantics, "epoch now", autograd._bw_epoch[0])
            for n, gv in zip(ar.layout.names, ar.grad_views):
                gskip[n.split("_", 1)[1]] = gv.detach().float().clone()
    wbefore = {n.split("_", 1)[1]: p.data().as_torch().detach().float().clone() for n, p in net.collect_params().items()}
    with autograd.record():
        l = lf(net(x1), y)
    l.backward()
    grads = {}
    for ar in tr._fused.arenas:
        for n, gv in zip(ar.layout.names, ar.grad_views):
            grads[n] = gv.detach().float().clone()
    for ar in tr._fused.arenas:
        print("after 3rd backward: dirty", ar.dirty, "bw_epoch", ar.bw_epoch, "epoch now", autograd._bw_epoch[0])
    run.extra = (gskip, wbefore)
    tr.step(8)
    return {n: p.data().as_torch().detach().float().clone() for n, p in net.collect_params().items() if p.grad_req != "null"}, grads


(a0, ga0) = run(False)
wb0 = run.extra[1]
(a, ga) = run(False)
wb1 = run.extra[1]
(b, gb) = run(True)
bad = []
for (n, u), v in zip(a.items(), b.values()):
    e = (u - v).abs().max().item() / (u.abs().max().item() + 1e-6)
    bad.append((e, n, tuple(u.shape), str(u.dtype)))
bad.sort(reverse=True)
print("worst weights:")
for r in bad[:8]:
    print("  %.4g %s %s" % (r[0], r[1], r[2]))
gbad = []
for (n, u), v in zip(ga.items(), gb.values()):
    e = (u - v).abs().max().item() / (u.abs().max().item() + 1e-6)
    gbad.append((e, n, tuple(u.shape)))
gbad.sort(reverse=True)
print("worst grads (arena, before step):")
for r in gbad[:12]:
    print("  %.4g %s %s" % r)
print("n grads differing > 1e-3:", sum(1 for r in gbad if r[0] > 1e-3), "of", len(gbad))

import torch
ga2 = {n.split("_", 1)[1]: v for n, v in ga.items()}
gb2 = {n.split("_", 1)[1]: v for n, v in gb.items()}
gskip, wb = run.extra
def nrm(d): return sum(float((v.double() ** 2).sum()) for v in d.values()) ** 0.5
print("||g_ref||", nrm(ga2), "||g_skiprun||", nrm(gb2), "||gskip||", nrm(gskip))
print("||g_skiprun - g_ref||", nrm({k: gb2[k] - ga2[k] for k in ga2}))
print("||g_skiprun - g_ref - gskip||", nrm({k: gb2[k] - ga2[k] - gskip[k] for k in ga2}))

ga02 = {n.split("_", 1)[1]: v for n, v in ga0.items()}
print("two identical runs: ||g_ref - g_ref'||", nrm({k: ga02[k] - ga2[k] for k in ga2}))
print("weights before 3rd backward: ref vs ref'", nrm({k: wb0[k] - wb1[k] for k in wb0}), " ref vs skiprun", nrm({k: wb1[k] - wb[k] for k in wb}), "||w||", nrm(wb))
big = sorted(((float((wb1[k] - wb[k]).abs().max()), k) for k in wb), reverse=True)[:6]
print("largest weight diffs before 3rd backward:", big)
