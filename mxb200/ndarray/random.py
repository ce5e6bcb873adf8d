"""mx.nd.random.* samplers (reference: src/operator/random/*, python/mxnet/ndarray/random.py)."""
import torch

from ..base import torch_dtype
from .ndarray import NDArray, _ctx_of, _shape

__all__ = ["uniform", "normal", "randn", "randint", "exponential", "gamma", "poisson", "multinomial", "shuffle", "seed"]


def _mk(shape, ctx, dtype):
    ctx = _ctx_of(ctx)
    shape = _shape(shape if shape is not None else (1,))
    return ctx, torch.empty(shape, dtype=torch_dtype(dtype), device=ctx.torch_device)


def _ret(t, ctx, out):
    if out is not None:
        out[:] = NDArray(t, ctx)
        return out
    return NDArray(t, ctx)


def uniform(low=0.0, high=1.0, shape=None, dtype=None, ctx=None, out=None, **kw):
    if out is not None and shape is None:
        shape, ctx = out.shape, out.context
    ctx, t = _mk(shape, ctx, dtype)
    return _ret(t.float().uniform_(low, high).to(t.dtype), ctx, out)


def normal(loc=0.0, scale=1.0, shape=None, dtype=None, ctx=None, out=None, **kw):
    if out is not None and shape is None:
        shape, ctx = out.shape, out.context
    ctx, t = _mk(shape, ctx, dtype)
    return _ret(t.float().normal_(loc, scale).to(t.dtype), ctx, out)


def randn(*shape, **kw):
    return normal(kw.pop("loc", 0.0), kw.pop("scale", 1.0), shape=shape, **kw)


def randint(low, high, shape=None, dtype="int32", ctx=None, out=None, **kw):
    ctx = _ctx_of(ctx)
    t = torch.randint(int(low), int(high), _shape(shape if shape is not None else (1,)), device=ctx.torch_device).to(torch_dtype(dtype))
    return _ret(t, ctx, out)


def exponential(scale=1.0, shape=None, dtype=None, ctx=None, out=None, **kw):
    ctx, t = _mk(shape, ctx, dtype)
    return _ret(t.float().exponential_(1.0 / scale).to(t.dtype), ctx, out)


def gamma(alpha=1.0, beta=1.0, shape=None, dtype=None, ctx=None, out=None, **kw):
    ctx, t = _mk(shape, ctx, dtype)
    d = torch.distributions.Gamma(torch.full_like(t.float(), alpha), torch.full_like(t.float(), 1.0 / beta))
    return _ret(d.sample().to(t.dtype), ctx, out)


def poisson(lam=1.0, shape=None, dtype=None, ctx=None, out=None, **kw):
    ctx, t = _mk(shape, ctx, dtype)
    return _ret(torch.poisson(torch.full_like(t.float(), lam)).to(t.dtype), ctx, out)


def multinomial(data, shape=1, get_prob=False, dtype="int32", **kw):
    n = shape if isinstance(shape, int) else int(torch.tensor(shape).prod())
    p = data._t.float()
    s = torch.multinomial(p.reshape(-1, p.shape[-1]), n, replacement=True)
    s = s.reshape(p.shape[:-1] + ((n,) if not (isinstance(shape, int) and shape == 1) else ()))
    return NDArray(s.to(torch_dtype(dtype)), data._ctx)


def shuffle(data, **kw):
    return NDArray(data._t[torch.randperm(data.shape[0], device=data._t.device)], data._ctx)


def seed(seed_state, ctx="all"):
    torch.manual_seed(int(seed_state))
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(int(seed_state))
