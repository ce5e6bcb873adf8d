"""LayerNorm over the last dim and fused bias+activation (csrc/cuda/layernorm.cu).

Reference: src/operator/nn/layer_norm-inl.h (forward = mean/std broadcast-reduce + 3 elementwise passes,
backward ≈ 8 generic launches); here one kernel per direction, fp32 statistics, dgamma/dbeta optionally
accumulated straight into an fp32 gradient-arena view (``sink_g`` / ``sink_b``)."""
import torch
import torch.nn.functional as F

from .. import _native as N
from ._dtypes import code as _dt
from .functional import LAUNCHES

_WS = {}
ACT_CODES = {"relu": 1, "gelu": 2, "gelu_erf": 2, "gelu_tanh": 3, "sigmoid": 4, "tanh": 5, "softrelu": 6, "leaky": 7,
             "elu": 8, "selu": 9, "silu": 10, "swish": 10}
_LN_WIDTHS = (256, 512, 768, 1024, 1536, 2048)


def _native_ok(x):
    return x.is_cuda and x.dtype in (torch.bfloat16, torch.float32) and N.have_cuda_kernels()


def _ws(dev, C):
    key = (dev, C)
    if key not in _WS:
        _WS[key] = torch.empty(N.cuda().MXBLayerNormWorkspaceFloats(N.c_int(C)), dtype=torch.float32, device=dev)
    return _WS[key]


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps, sink_g, sink_b, residual):
        C = x.shape[-1]
        x2 = x.contiguous().view(-1, C)
        rows = x2.shape[0]
        y = torch.empty_like(x2)
        r2 = residual.contiguous().view(-1, C) if residual is not None else None
        xs = torch.empty_like(x2) if r2 is not None else None
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        g32 = gamma.float().contiguous() if gamma is not None else None
        b32 = beta.float().contiguous() if beta is not None else None
        lib = N.cuda()
        N.check(lib, lib.MXBLayerNormForward(N.ptr(x2), N.ptr(r2), N.ptr(xs), N.c_int(_dt(x.dtype)), N.ptr(g32), N.ptr(b32),
                                             N.ptr(y), N.ptr(mean), N.ptr(rstd), N.c_int64(rows), N.c_int(C),
                                             N.c_float(eps), N.stream_of(x2)))
        LAUNCHES[0] += 1
        ctx.has_res = r2 is not None
        ctx.save_for_backward(xs if r2 is not None else x2, g32, mean, rstd)
        ctx.shape = x.shape
        ctx.has = (gamma is not None, beta is not None)
        ctx.gdt = (gamma.dtype if gamma is not None else None, beta.dtype if beta is not None else None)
        ctx.sinks = (sink_g, sink_b)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, g32, mean, rstd = ctx.saved_tensors
        C = x2.shape[-1]
        rows = x2.shape[0]
        dy2 = dy.contiguous().view(-1, C)
        dx = torch.empty_like(x2)
        sink_g, sink_b = ctx.sinks
        from .functional import sink_touch
        sink_touch(sink_g, sink_b)
        use_sink = sink_g is not None and sink_b is not None
        if use_sink:
            dg, db = sink_g, sink_b
        else:
            dg = torch.empty(C, dtype=torch.float32, device=x2.device) if ctx.has[0] else None
            db = torch.empty(C, dtype=torch.float32, device=x2.device) if ctx.has[1] else None
        lib = N.cuda()
        N.check(lib, lib.MXBLayerNormBackward(N.ptr(dy2), N.ptr(x2), N.c_int(_dt(x2.dtype)), N.ptr(g32), N.ptr(mean),
                                              N.ptr(rstd), N.ptr(dx), N.ptr(dg), N.ptr(db), N.ptr(_ws(x2.device, C)),
                                              N.c_int64(rows), N.c_int(C), N.c_int(1 if use_sink else 0),
                                              N.stream_of(x2)))
        LAUNCHES[0] += 1
        dxv = dx.view(ctx.shape)
        dres = dxv if ctx.has_res else None           # d(x + r) flows unchanged to both addends
        if use_sink:
            return dxv, None, None, None, None, None, dres
        return (dxv, dg.to(ctx.gdt[0]) if dg is not None else None,
                db.to(ctx.gdt[1]) if db is not None else None, None, None, None, dres)


def layer_norm(x, gamma, beta, eps=1e-5, sink_g=None, sink_b=None, residual=None):
    """LayerNorm over the last dim; ``residual``: y = LN(x + residual) with the add fused into the same pass."""
    C = x.shape[-1]
    if _native_ok(x) and C in _LN_WIDTHS and (residual is None or residual.dtype is x.dtype):
        return _LayerNormFn.apply(x, gamma, beta, float(eps), sink_g, sink_b, residual)
    if residual is not None:
        x = x + residual
    return F.layer_norm(x, (C,), gamma.to(x.dtype) if gamma is not None else None,
                        beta.to(x.dtype) if beta is not None else None, eps)


class _BiasActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias, act, slope, sink_b):
        C = x.shape[-1]
        x2 = x.contiguous().view(-1, C)
        b32 = bias.float().contiguous() if bias is not None else None
        y = torch.empty_like(x2)
        lib = N.cuda()
        N.check(lib, lib.MXBBiasActForward(N.ptr(x2), N.c_int(_dt(x.dtype)), N.ptr(b32), N.ptr(y), N.c_int64(x2.shape[0]),
                                           N.c_int(C), N.c_int(act), N.c_float(slope), N.stream_of(x2)))
        LAUNCHES[0] += 1
        ctx.save_for_backward(x2, b32)
        ctx.meta = (x.shape, act, slope, bias.dtype if bias is not None else None, sink_b)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, b32 = ctx.saved_tensors
        shape, act, slope, bdt, sink_b = ctx.meta
        from .functional import sink_touch
        sink_touch(sink_b)
        C = x2.shape[-1]
        dy2 = dy.contiguous().view(-1, C)
        dx = torch.empty_like(x2)
        lib = N.cuda()
        N.check(lib, lib.MXBBiasActBackward(N.ptr(dy2), N.ptr(x2), N.c_int(_dt(x2.dtype)), N.ptr(b32), N.ptr(dx),
                                            N.c_int64(x2.shape[0]), N.c_int(C), N.c_int(act), N.c_float(slope),
                                            N.stream_of(x2)))
        LAUNCHES[0] += 1
        db = None
        if b32 is not None:
            from .functional import colsum
            if sink_b is not None:       # fp32 gradient-arena view: accumulate there, autograd gets nothing
                colsum(dx, out=sink_b)
            else:
                db = colsum(dx).to(bdt)
        return dx.view(shape), db, None, None, None


def _torch_act(x, act, slope):
    if act == "relu":
        return torch.relu(x)
    if act in ("gelu", "gelu_erf"):
        return F.gelu(x)
    if act == "gelu_tanh":
        return F.gelu(x, approximate="tanh")
    if act == "sigmoid":
        return torch.sigmoid(x)
    if act == "tanh":
        return torch.tanh(x)
    if act == "softrelu":
        return F.softplus(x)
    if act == "leaky":
        return F.leaky_relu(x, slope)
    if act == "elu":
        return F.elu(x, slope)
    if act == "selu":
        return F.selu(x)
    if act in ("silu", "swish"):
        return F.silu(x)
    raise ValueError("unknown activation %r" % (act,))


def bias_act(x, bias=None, act="gelu", slope=0.25, sink_b=None):
    """y = act(x + bias) along the last dim in one pass (and one pass backward)."""
    if act not in ACT_CODES:
        raise ValueError("unknown activation %r" % (act,))
    C = x.shape[-1]
    if _native_ok(x) and C % 8 == 0 and x.numel() > 0:
        return _BiasActFn.apply(x, bias, ACT_CODES[act], float(slope), sink_b)
    return _torch_act(x + bias.to(x.dtype) if bias is not None else x, act, slope)
