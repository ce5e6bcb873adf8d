"""Forward the fused ResNet-18 twice on the same input (training mode) and find the first op whose output differs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mxb200 as mx
from mxb200 import autograd
from mxb200.gluon.model_zoo import vision
from mxb200.ops import functional as K

torch.manual_seed(0)
mx.random.seed(0)
ctx = mx.gpu(0)
net = vision.get_resnet("v1b", 18, classes=10, thumbnail=True, ctx=ctx, dtype="bfloat16", layout="NHWC")
net.hybridize()
g = torch.Generator(device="cuda").manual_seed(3)
x0 = mx.nd.from_torch(torch.randn(8, 32, 32, 3, device="cuda", generator=g).bfloat16(), ctx)

trace = []
for name in ["conv_nhwc", "conv1x1_nhwc", "stem_conv_nhwc", "bn_act", "linear", "max_pool_nhwc"]:
    orig = getattr(K, name)

    def wrap(*a, _o=orig, _n=name, **kw):
        r = _o(*a, **kw)
        t = r[0] if isinstance(r, (tuple, list)) else r
        ins = a[0]
        trace.append((_n, tuple(t.shape), t.detach().float().clone(), ins.detach().float().clone() if torch.is_tensor(ins) else None))
        return r
    setattr(K, name, wrap)

runs = []
for i in range(2):
    trace.clear()
    with autograd.record():
        out = net(x0)
    torch.cuda.synchronize()
    runs.append(list(trace))
print("ops traced:", len(runs[0]))
first = True
for i, ((n1, s1, t1, i1), (n2, s2, t2, i2)) in enumerate(zip(*runs)):
    d = (t1 - t2).abs().max().item()
    di = (i1 - i2).abs().max().item() if i1 is not None else -1
    if d > 0 or di > 0:
        print("op %3d %-16s out %-20s out diff %.4g (max %.4g)   input diff %.4g" % (i, n1, s1, d, t1.abs().max().item(), di))
        if first and di == 0:
            first = False
            bad = (t1 - t2).abs()
            idx = bad.flatten().topk(5).indices
            print("     >>> FIRST divergence with identical input; #elements differing:", int((bad > 0).sum()), "of", bad.numel())
