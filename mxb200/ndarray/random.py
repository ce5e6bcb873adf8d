"""mx.nd.random.* samplers (reference: src/operator/random/*, python/mxnet/ndarray/random.py)."""
import torch

from ..base import torch_dtype
from .ndarray import NDArray, _ctx_of, _shape

__all__ = ["uniform", "normal", "randn", "randint", "exponential", "gamma", "poisson", "negative_binomial",
           "generalized_negative_binomial", "multinomial", "shuffle", "seed"]


def _mk(shape, ctx, dtype):
    ctx = _ctx_of(ctx)
    shape = _shape(shape if shape is not None else (1,))
    return ctx, torch.empty(shape, dtype=torch_dtype(dtype), device=ctx.torch_device)


def _ret(t, ctx, out):
    if out is not None:
        out[:] = NDArray(t, ctx)
        return out
    return NDArray(t, ctx)


def _random_helper(name, params, shape, dtype, ctx, out, kwargs):
    """Scalar parameters -> ``_random_<name>`` (one distribution, ``shape`` samples); NDArray parameters -> ``_sample_<name>``
    (one distribution per element, output shape = params.shape + shape) -- python/mxnet/ndarray/random.py:31-51."""
    from .op import _REG
    if all(isinstance(p, NDArray) for p in params):
        return _REG["_sample_" + name](*params, shape=shape, dtype=dtype, out=out, **kwargs)
    if any(isinstance(p, NDArray) for p in params):
        raise ValueError("Distribution parameters must all have the same type, but got both %s" % str([type(p) for p in params]))
    if out is not None and shape is None:
        shape, ctx = out.shape, out.context
    if shape is None:
        shape = (1,)
    return _REG["_random_" + name](*params, shape=shape, dtype=dtype, ctx=ctx, out=out, **kwargs)


def uniform(low=0, high=1, shape=None, dtype=None, ctx=None, out=None, **kwargs):
    return _random_helper("uniform", [low, high], shape, dtype, ctx, out, kwargs)


def normal(loc=0, scale=1, shape=None, dtype=None, ctx=None, out=None, **kwargs):
    if not isinstance(scale, NDArray) and scale < 0:
        # invalid parameters fail like any asynchronous operator error: at the next synchronisation point
        from ..base import MXNetError
        from .ndarray import _poison
        res = _random_helper("normal", [loc, 0.0], shape, dtype, ctx, out, kwargs)
        _poison(res._t, MXNetError("normal: scale (standard deviation) must be non-negative, got %s" % scale))
        return res
    return _random_helper("normal", [loc, scale], shape, dtype, ctx, out, kwargs)


def randn(*shape, **kwargs):
    loc, scale = kwargs.pop("loc", 0), kwargs.pop("scale", 1)
    return normal(loc, scale, shape=shape or None, **kwargs)


def randint(low, high, shape=None, dtype="int32", ctx=None, out=None, **kw):
    ctx = _ctx_of(ctx)
    t = torch.randint(int(low), int(high), _shape(shape if shape is not None else (1,)), device=ctx.torch_device).to(torch_dtype(dtype))
    return _ret(t, ctx, out)


def poisson(lam=1, shape=None, dtype=None, ctx=None, out=None, **kwargs):
    return _random_helper("poisson", [lam], shape, dtype, ctx, out, kwargs)


def exponential(scale=1, shape=None, dtype=None, ctx=None, out=None, **kwargs):
    # the operators are parameterised by the rate lam = 1/scale (sample_op.h ExponentialParam)
    return _random_helper("exponential", [1.0 / scale], shape, dtype, ctx, out, kwargs)


def gamma(alpha=1, beta=1, shape=None, dtype=None, ctx=None, out=None, **kwargs):
    return _random_helper("gamma", [alpha, beta], shape, dtype, ctx, out, kwargs)


def negative_binomial(k=1, p=1, shape=None, dtype=None, ctx=None, out=None, **kwargs):
    return _random_helper("negative_binomial", [k, p], shape, dtype, ctx, out, kwargs)


def generalized_negative_binomial(mu=1, alpha=1, shape=None, dtype=None, ctx=None, out=None, **kwargs):
    return _random_helper("generalized_negative_binomial", [mu, alpha], shape, dtype, ctx, out, kwargs)


def multinomial(data, shape=None, get_prob=False, out=None, dtype="int32", **kwargs):
    """Samples category indices from the distributions along the last axis of ``data``; with ``get_prob`` also returns the
    log-likelihood of the samples (for REINFORCE-style gradients)."""
    from .op import _REG
    return _REG["_sample_multinomial"](data, shape=shape, get_prob=get_prob, dtype=dtype, out=out)


def shuffle(data, **kwargs):
    from .op import _REG
    return _REG["_shuffle"](data, **kwargs)


def seed(seed_state, ctx="all"):
    torch.manual_seed(int(seed_state))
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(int(seed_state))
