"""Run every hot op twice on identical inputs (ResNet-18-thumbnail shapes, batch 8) and report the largest difference
between the two results: outputs written without atomics must be bit-identical."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mxb200.ops import functional as K

dev = torch.device("cuda", 0)
torch.manual_seed(0)


def cmp(name, a, b):
    if not isinstance(a, (list, tuple)):
        a, b = [a], [b]
    out = []
    for u, v in zip(a, b):
        if u is None:
            continue
        d = (u.float() - v.float()).abs().max().item()
        out.append("%.3g/%.3g" % (d, u.float().abs().max().item()))
    print("%-46s max|run1-run2| / max|x| : %s" % (name, "  ".join(out)), flush=True)


def twice(fn):
    r1 = fn()
    torch.cuda.synchronize()
    r2 = fn()
    torch.cuda.synchronize()
    return r1, r2


for (n, h, c, o, k, s) in [(8, 32, 64, 128, 1, 2),
                           (8, 16, 128, 256, 1, 2), (8, 8, 256, 512, 1, 2)]:
    x = torch.randn(n, h, h, c, device=dev).bfloat16()
    w = (torch.randn(o, c, k, k, device=dev) * 0.05).bfloat16().contiguous(memory_format=torch.channels_last)
    pad = k // 2
    P = (h + 2 * pad - k) // s + 1
    dy = torch.randn(n, P, P, o, device=dev).bfloat16()

    def fwd():
        ws = K.bn_stats_workspace_for(dev, o)
        if ws is not None:
            ws.zero_()
        if k == 1:
            y = K.conv1x1_nhwc(x, w.view(o, c), s, stats=ws)
        else:
            y = K.conv_nhwc(x, w, s, pad, stats=ws)
        st = ws[:2 * o].clone() if ws is not None else None
        if ws is not None:
            ws.zero_()
        return [y, st]

    cmp("conv%dx%d fwd+stats n%d %dx%d %d->%d s%d" % (k, k, n, h, h, c, o, s), *twice(fwd))

    def bwd():
        xx = x.clone().requires_grad_(True)
        ww = (w.view(o, c) if k == 1 else w).clone().requires_grad_(True)
        y = K.conv1x1_nhwc(xx, ww, s) if k == 1 else K.conv_nhwc(xx, ww, s, pad)
        y.backward(dy)
        return [xx.grad, ww.grad]

    cmp("conv%dx%d bwd (dx, dw)" % (k, k), *twice(bwd))

for (M, C, res) in [(8 * 32 * 32, 64, False), (8 * 32 * 32, 64, True), (8 * 16 * 16, 128, True), (8 * 8 * 8, 256, True), (8 * 4 * 4, 512, True),
                    (8 * 4 * 4, 512, False)]:
    x = torch.randn(M, C, device=dev).bfloat16()
    r = torch.randn(M, C, device=dev).bfloat16() if res else None
    dy = torch.randn(M, C, device=dev).bfloat16()

    def bn():
        xx = x.clone().requires_grad_(True)
        rr = r.clone().requires_grad_(True) if res else None
        g = torch.ones(C, device=dev, requires_grad=True)
        b = torch.zeros(C, device=dev, requires_grad=True)
        rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
        y = K.bn_act(xx, g, b, rm, rv, residual=rr, relu=True)
        y.backward(dy)
        return [y.detach(), xx.grad, g.grad, b.grad, rr.grad if res else None, rm, rv]

    cmp("bn_act fwd/bwd M=%d C=%d res=%d (y dx dg db dr rm rv)" % (M, C, res), *twice(bn))

x = torch.randn(8, 512, device=dev).bfloat16()
w = (torch.randn(10, 512, device=dev) * 0.05).bfloat16()
b = torch.zeros(10, device=dev).bfloat16()


def fc():
    xx, ww, bb = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = K.linear(xx, ww, bb)
    y.backward(torch.ones_like(y))
    return [y.detach(), xx.grad, ww.grad, bb.grad]


cmp("linear 8x512 -> 10", *twice(fc))
x = torch.randn(8, 32, 32, 3, device=dev).bfloat16()
w = (torch.randn(64, 3, 3, 3, device=dev) * 0.1).bfloat16()


def stem():
    ww = w.clone().requires_grad_(True)
    y = K.stem_conv_nhwc(x, ww, 1, 1)
    y.backward(torch.ones_like(y))
    return [y.detach(), ww.grad]


cmp("stem conv 3->64", *twice(stem))
