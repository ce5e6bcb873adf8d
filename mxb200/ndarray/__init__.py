"""``mx.nd`` namespace."""
import sys as _sys
import types as _types

from .ndarray import from_dlpack, to_dlpack_for_read, to_dlpack_for_write  # noqa: F401
from .ndarray import (NDArray, array, zeros, ones, empty, full, arange, zeros_like, ones_like, from_torch, waitall,
                      concatenate, moveaxis)
from . import op, sparse, random, serialization
from .serialization import save, load, load_frombuffer
from .sparse import RowSparseNDArray, CSRNDArray, BaseSparseNDArray, cast_storage
from .op import *  # noqa: F401,F403
from . import op_extra as _op_extra  # noqa: F401  (registers the long tail into _REG)
from . import op_legacy as _op_legacy  # noqa: F401  (legacy aliases + deformable ops)
from .op import _REG as _OPS
from .op import CachedOp  # noqa: F401

import functools as _functools


def _strip_sym_kwargs(f, opname=None):
    """Ops accept (and ignore) the symbolic-only ``name=`` / ``attr=`` kwargs so HybridBlocks can pass them.  Operators run
    on the dense view of sparse inputs (storage fallback); ``sparse.storage_rule`` then gives the result the storage type
    the reference's FInferStorageType would have picked (relu(row_sparse) stays row_sparse ...)."""
    @_functools.wraps(f)
    def g(*a, **k):
        k.pop("name", None)
        k.pop("attr", None)
        r = f(*a, **k)
        for x in a:
            if isinstance(x, BaseSparseNDArray):
                return sparse.storage_rule(opname, a, k, r)
        return r
    return g


_this = _sys.modules[__name__]
for _k in list(_OPS.keys()):
    _OPS[_k] = _strip_sym_kwargs(_OPS[_k], _k)
    setattr(_this, _k, _OPS[_k])
# creation functions win over same-named elementwise ops
zeros_like = ndarray.zeros_like
ones_like = ndarray.ones_like

# mx.nd.contrib.* : ops registered as _contrib_X
contrib = _types.ModuleType(__name__ + ".contrib")
for _k, _v in _OPS.items():
    if _k.startswith("_contrib_"):
        setattr(contrib, _k[len("_contrib_"):], _v)
_sys.modules[contrib.__name__] = contrib

def _rand_zipfian(true_classes, num_sampled, range_max, ctx=None):
    """Log-uniform (Zipfian) candidate sampler for sampled-softmax style losses (ndarray/contrib.py:rand_zipfian of the
    reference): P(class) = log((class + 2) / (class + 1)) / log(range_max + 1).  Returns (sampled classes int64,
    expected count of the true classes, expected count of the sampled classes), both counts float64."""
    import math as _math
    ctx = ctx if ctx is not None else true_classes.context
    log_range = _math.log(range_max + 1)
    r = random.uniform(0, log_range, shape=(num_sampled,), dtype="float64", ctx=ctx)
    sampled = NDArray(((r._t.exp() - 1).to(_torch.int64)) % range_max, ctx)
    def _expected(cls):
        c = cls._t.to(_torch.float64)
        return NDArray(((c + 2.0) / (c + 1.0)).log() / log_range * num_sampled, ctx)
    return sampled, _expected(true_classes.as_in_context(ctx)), _expected(sampled)


import torch as _torch  # noqa: E402
contrib.rand_zipfian = _rand_zipfian

linalg = _types.ModuleType(__name__ + ".linalg")
for _k, _v in _OPS.items():
    if _k.startswith("linalg_"):
        setattr(linalg, _k[len("linalg_"):], _v)
_sys.modules[linalg.__name__] = linalg


# mx.nd._internal / mx.nd.op / mx.nd.image: the same registry under the reference's other module names
class _Prefixed(_types.ModuleType):
    def __init__(self, name, prefixes, fallback=True):
        super(_Prefixed, self).__init__(name)
        self._prefixes, self._fallback = prefixes, fallback

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        for pre in self._prefixes:
            if pre + name in _OPS:
                return _OPS[pre + name]
        if self._fallback and name in _OPS:
            return _OPS[name]
        raise AttributeError("module %r has no attribute %r" % (self.__name__, name))


_internal = _Prefixed(__name__ + "._internal", ("",))
image = _Prefixed(__name__ + ".image", ("_image_",), fallback=False)
_sys.modules[_internal.__name__] = _internal
_sys.modules[image.__name__] = image
for _k in ("negative_binomial", "generalized_negative_binomial"):
    pass
def eye(N, M=0, k=0, ctx=None, dtype=None, **kwargs):
    """2-D array with ones on the k-th diagonal (ndarray.py:eye of the reference)."""
    return _OPS["_eye"](N=N, M=M, k=k, ctx=ctx, dtype=dtype or "float32", **kwargs)


# legacy top-level sampler names (mx.nd.uniform / mx.nd.normal, python/mxnet/ndarray/random.py re-exports)
uniform, normal = random.uniform, random.normal


# ----------------------------------------------------------------------------- fluent methods (x.op(...) == mx.nd.op(x, ...))
# The reference generates one NDArray method per "fluent" operator (python/mxnet/ndarray/ndarray.py:1019-1500).  The
# hand-written methods of NDArray stay; reductions are re-pointed at the operators so they take the full argument set
# (exclude=, out=), and every other reference method that is still missing is added from the operator registry.
_FLUENT_FORCE = ("sum", "mean", "prod", "max", "min", "nansum", "nanprod", "argmax", "argmin", "norm", "topk", "sort", "argsort",
                 "clip", "broadcast_like", "flip", "tile", "repeat", "pad", "one_hot", "pick", "take", "slice", "slice_like")
_FLUENT_ADD = ("arccos", "arccosh", "arcsin", "arcsinh", "arctan", "arctanh", "argmax_channel", "broadcast_axes", "cbrt",
               "ceil", "cos", "cosh", "degrees", "depth_to_space", "diag", "expm1", "fix", "floor", "log10", "log1p", "log2",
               "radians", "rcbrt", "reciprocal", "rint", "round", "rsqrt", "shape_array", "sin", "sinh", "size_array", "softmin",
               "space_to_depth", "tan", "trunc", "sign", "softmax", "log_softmax", "zeros_like", "ones_like", "reshape_like",
               "expand_dims", "squeeze", "relu", "sigmoid", "tanh", "exp", "log", "sqrt", "square", "abs", "flatten", "swapaxes",
               "split", "slice_axis", "broadcast_to", "transpose")


def _install_fluent():
    def make(fn, name):
        def method(self, *args, **kwargs):
            return fn(self, *args, **kwargs)
        method.__name__ = name
        method.__doc__ = "Convenience fluent method for :py:func:`%s`." % name
        return method
    for _n in _FLUENT_FORCE:
        if _n in _OPS:
            setattr(NDArray, _n, make(_OPS[_n], _n))
    for _n in _FLUENT_ADD:
        if _n in _OPS and not hasattr(NDArray, _n):
            setattr(NDArray, _n, make(_OPS[_n], _n))


_install_fluent()


def __getattr__(name):
    """``mx.nd.Custom`` before ``mx.operator`` was touched: importing the operator module registers it."""
    if name == "Custom":
        import importlib
        importlib.import_module(__name__.rsplit(".", 1)[0] + ".operator")
        return _OPS["Custom"]
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
