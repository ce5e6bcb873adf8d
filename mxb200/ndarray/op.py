"""Operator namespace ``mx.nd.*`` (reference: src/operator/** registered through NNVM and
exposed in python/mxnet/ndarray/op.py).

Each op is a function over NDArrays implemented with torch; the hot ops
(FullyConnected, Convolution+BatchNorm+ReLU, SoftmaxOutput, optimizer
updates, LayerNorm) dispatch to the hand-written sm_100a kernels in
``mxb200.ops`` when their inputs are CUDA tensors.  Semantics follow
SURVEY.md Appendix C (file:line cites inline).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from ..base import MXNetError, torch_dtype
from .. import autograd as _ag
from .ndarray import NDArray, _op, _unwrap, _wrap, _result_ctx, _mx_reshape

_REG = {}


def register(fn=None, name=None):
    def deco(f):
        _REG[name or f.__name__] = f
        return f
    return deco(fn) if fn is not None else deco


def _out(res, out):
    if out is None:
        return res
    if isinstance(out, (list, tuple)):
        for o, r in zip(out, res):
            o[:] = r
        return out
    out[:] = res
    return out


def _axes(axis):
    if axis is None or axis == ():
        return None
    return tuple(axis) if isinstance(axis, (list, tuple)) else int(axis)


# ---------------------------------------------------------------------------
# elementwise / broadcast
# ---------------------------------------------------------------------------

def _unary(name, fn):
    def f(data, out=None, **kw):
        return _out(_op(fn, data), out)
    f.__name__ = name
    _REG[name] = f
    return f


abs = _unary("abs", torch.abs)
sign = _unary("sign", torch.sign)
sqrt = _unary("sqrt", torch.sqrt)
rsqrt = _unary("rsqrt", torch.rsqrt)
square = _unary("square", torch.square)
exp = _unary("exp", torch.exp)
expm1 = _unary("expm1", torch.expm1)
log = _unary("log", torch.log)
log2 = _unary("log2", torch.log2)
log10 = _unary("log10", torch.log10)
log1p = _unary("log1p", torch.log1p)
sin = _unary("sin", torch.sin)
cos = _unary("cos", torch.cos)
tan = _unary("tan", torch.tan)
arcsin = _unary("arcsin", torch.asin)
arccos = _unary("arccos", torch.acos)
arctan = _unary("arctan", torch.atan)
sinh = _unary("sinh", torch.sinh)
cosh = _unary("cosh", torch.cosh)
tanh = _unary("tanh", torch.tanh)
sigmoid = _unary("sigmoid", torch.sigmoid)
relu = _unary("relu", torch.relu)
floor = _unary("floor", torch.floor)
ceil = _unary("ceil", torch.ceil)
round = _unary("round", torch.round)
trunc = _unary("trunc", torch.trunc)
negative = _unary("negative", torch.neg)
reciprocal = _unary("reciprocal", torch.reciprocal)
erf = _unary("erf", torch.erf)
gamma = _unary("gamma", lambda t: torch.exp(torch.lgamma(t)))
gammaln = _unary("gammaln", torch.lgamma)
softsign = _unary("softsign", F.softsign)
identity = _unary("identity", lambda t: t.clone())
stop_gradient = _unary("stop_gradient", lambda t: t.detach())
BlockGrad = stop_gradient
_REG["BlockGrad"] = BlockGrad
zeros_like = _unary("zeros_like", torch.zeros_like)
ones_like = _unary("ones_like", torch.ones_like)
logical_not = _unary("logical_not", lambda t: (t == 0).to(t.dtype))


def _binary(name, fn):
    def f(lhs, rhs, out=None, **kw):
        if not isinstance(lhs, NDArray):
            return _out(_op(lambda b: fn(lhs, b), rhs), out)
        if not isinstance(rhs, NDArray):
            return _out(_op(lambda a: fn(a, rhs), lhs), out)
        return _out(_op(fn, lhs, rhs), out)
    f.__name__ = name
    _REG[name] = f
    return f


def _cmpf(fn):
    return lambda a, b: fn(a, b).to(a.dtype if isinstance(a, torch.Tensor) else b.dtype)


add = _binary("add", torch.add)
subtract = _binary("subtract", torch.sub)
multiply = _binary("multiply", torch.mul)
divide = _binary("divide", torch.div)
def _mx_mod(a, b):
    """mshadow_op::mod: python-style remainder, and x mod 0 = 0 (also for integers, where torch raises)."""
    ta = a if isinstance(a, torch.Tensor) else b
    if not ta.is_floating_point():
        bt = b if isinstance(b, torch.Tensor) else torch.as_tensor(b, device=ta.device)
        zero = bt == 0
        r = torch.remainder(a, torch.where(zero, torch.ones_like(bt), bt))
        return torch.where(zero, torch.zeros_like(r), r)
    r = torch.remainder(a, b)
    return torch.where(torch.as_tensor(b, device=r.device) == 0, torch.zeros_like(r), r) if isinstance(b, torch.Tensor) else (r if b != 0 else torch.zeros_like(r))


modulo = _binary("modulo", _mx_mod)
power = _binary("power", torch.pow)
def _mx_maximum(a, b):
    """Element-wise maximum of two arrays with the reference's tie rule for the gradient: at a == b the LEFT operand takes
    all of it (mshadow_op ge / lt), where torch.maximum splits it in halves."""
    return torch.where(a >= b, a, b)


def _mx_minimum(a, b):
    return torch.where(a <= b, a, b)


maximum = _binary("maximum", lambda a, b: _mx_maximum(a, b) if isinstance(b, torch.Tensor) and isinstance(a, torch.Tensor)
                  else torch.clamp(a if isinstance(a, torch.Tensor) else b, min=b if isinstance(a, torch.Tensor) else a))
minimum = _binary("minimum", lambda a, b: _mx_minimum(a, b) if isinstance(b, torch.Tensor) and isinstance(a, torch.Tensor)
                  else torch.clamp(a if isinstance(a, torch.Tensor) else b, max=b if isinstance(a, torch.Tensor) else a))
hypot = _binary("hypot", torch.hypot)
equal = _binary("equal", _cmpf(torch.eq))
not_equal = _binary("not_equal", _cmpf(torch.ne))
greater = _binary("greater", _cmpf(torch.gt))
greater_equal = _binary("greater_equal", _cmpf(torch.ge))
lesser = _binary("lesser", _cmpf(torch.lt))
lesser_equal = _binary("lesser_equal", _cmpf(torch.le))
logical_and = _binary("logical_and", lambda a, b: ((a != 0) & (b != 0)).to(a.dtype))
logical_or = _binary("logical_or", lambda a, b: ((a != 0) | (b != 0)).to(a.dtype))
logical_xor = _binary("logical_xor", lambda a, b: ((a != 0) ^ (b != 0)).to(a.dtype))
for _n, _f in [("broadcast_add", add), ("broadcast_plus", add), ("broadcast_sub", subtract), ("broadcast_minus", subtract),
               ("broadcast_mul", multiply), ("broadcast_div", divide), ("broadcast_mod", modulo),
               ("broadcast_power", power), ("broadcast_maximum", maximum), ("broadcast_minimum", minimum),
               ("broadcast_hypot", hypot), ("broadcast_equal", equal), ("broadcast_not_equal", not_equal),
               ("broadcast_greater", greater), ("broadcast_greater_equal", greater_equal),
               ("broadcast_lesser", lesser), ("broadcast_lesser_equal", lesser_equal),
               ("broadcast_logical_and", logical_and), ("broadcast_logical_or", logical_or),
               ("broadcast_logical_xor", logical_xor), ("elemwise_add", add), ("elemwise_sub", subtract),
               ("elemwise_mul", multiply), ("elemwise_div", divide)]:
    _REG[_n] = _f
    globals()[_n] = _f


@register
def clip(data, a_min, a_max, out=None):
    return _out(_op(lambda t: torch.clamp(t, a_min, a_max), data), out)


@register
def add_n(*args, out=None):
    """elemwise_sum (src/operator/tensor/elemwise_sum.cu)."""
    return _out(_op(lambda *ts: torch.stack(ts).sum(0) if len(ts) > 4 else _sum_small(ts), *args), out)


def _sum_small(ts):
    r = ts[0] + ts[1] if len(ts) > 1 else ts[0].clone()
    for t in ts[2:]:
        r = r + t
    return r


ElementWiseSum = add_n
_REG["ElementWiseSum"] = add_n


@register
def where(condition, x, y, out=None):
    """x where condition != 0 else y; a 1-D condition selects whole rows (control_flow_op.h WhereOpForward)."""
    def g(c, a, b):
        if c.dim() == 1 and a.dim() > 1:
            c = c.reshape((-1,) + (1,) * (a.dim() - 1))
        return torch.where(c != 0, a, b)
    return _out(_op(g, condition, x, y), out)


@register
def cast(data, dtype, out=None):
    return _out(_op(lambda t: t.to(torch_dtype(dtype)), data), out)


Cast = cast
_REG["Cast"] = cast


@register
def smooth_l1(data, scalar=1.0, out=None):
    s2 = scalar * scalar
    return _out(_op(lambda t: torch.where(t.abs() < 1.0 / s2, 0.5 * s2 * t * t, t.abs() - 0.5 / s2), data), out)


# ---------------------------------------------------------------------------
# reductions
# ---------------------------------------------------------------------------

def _reduce(name, fn):
    def f(data, axis=None, keepdims=False, exclude=False, out=None, **kw):
        ax = _axes(axis)
        if exclude and ax is not None:
            axs = (ax,) if isinstance(ax, int) else ax
            axs = tuple(a % data.ndim for a in axs)
            ax = tuple(i for i in range(data.ndim) if i not in axs)

        def g(t):
            if ax is None:
                r = fn(t)
                return r.reshape((1,) * t.dim()) if keepdims else r.reshape(1)
            if isinstance(ax, tuple) and fn in (torch.prod,):      # torch.prod takes one axis at a time
                r = t
                for a in sorted((a % t.dim() for a in ax), reverse=True):
                    r = fn(r, dim=a, keepdim=bool(keepdims))
                return r.reshape(1) if r.dim() == 0 else r
            r = fn(t, dim=ax, keepdim=bool(keepdims))
            return r.reshape(1) if r.dim() == 0 else r           # a full reduction has shape (1,) in MXNet 1.x
        return _out(_op(g, data), out)
    f.__name__ = name
    _REG[name] = f
    return f


sum = _reduce("sum", torch.sum)
mean = _reduce("mean", torch.mean)
prod = _reduce("prod", torch.prod)
max = _reduce("max", torch.amax)
min = _reduce("min", torch.amin)
nansum = _reduce("nansum", torch.nansum)
_REG["sum_axis"] = sum
sum_axis = sum


@register
def norm(data, ord=2, axis=None, keepdims=False, out=None):
    def g(t):
        r = torch.linalg.vector_norm(t.float(), ord=ord, dim=_axes(axis), keepdim=bool(keepdims)).to(t.dtype)
        return r.reshape(1) if r.dim() == 0 else r
    return _out(_op(g, data), out)


@register
def argmax(data, axis=None, keepdims=False, out=None):
    return _out(_op(lambda t: torch.argmax(t, dim=axis, keepdim=bool(keepdims)).to(torch.float32), data), out)


@register
def argmin(data, axis=None, keepdims=False, out=None):
    return _out(_op(lambda t: torch.argmin(t, dim=axis, keepdim=bool(keepdims)).to(torch.float32), data), out)


@register
def argmax_channel(data, out=None):
    return argmax(data, axis=1, out=out)


@register
def topk(data, axis=-1, k=1, ret_typ="indices", is_ascend=False, dtype="float32", out=None):
    if axis == "None":
        axis = None

    def g(t):
        if axis is None:
            t = t.reshape(-1)
        v, i = torch.topk(t, k, dim=-1 if axis is None else axis, largest=not is_ascend)
        i = i.to(torch_dtype(dtype))
        if ret_typ == "indices":
            return i
        if ret_typ == "value":
            return v
        if ret_typ == "both":
            return [v, i]
        m = torch.zeros_like(t)
        m.scatter_(-1 if axis is None else axis, i.long(), 1)
        return m.reshape(data.shape) if axis is None else m
    return _out(_op(g, data), out)


@register
def sort(data, axis=-1, is_ascend=True, out=None):
    if axis is None or axis == "None":           # the flattened array is sorted (ordering_op-inl.h: axis=None)
        return _out(_op(lambda t: torch.sort(t.reshape(-1), dim=0, descending=not is_ascend)[0], data), out)
    return _out(_op(lambda t: torch.sort(t, dim=axis, descending=not is_ascend)[0], data), out)


@register
def argsort(data, axis=-1, is_ascend=True, dtype="float32", out=None):
    if axis is None or axis == "None":
        return _out(_op(lambda t: torch.argsort(t.reshape(-1), dim=0, descending=not is_ascend).to(torch_dtype(dtype)), data), out)
    return _out(_op(lambda t: torch.argsort(t, dim=axis, descending=not is_ascend).to(torch_dtype(dtype)), data), out)


# ---------------------------------------------------------------------------
# shape manipulation
# ---------------------------------------------------------------------------

@register
def reshape(data, shape=None, reverse=False, target_shape=None, keep_highest=False, out=None):
    if shape is None or len(shape) == 0:
        # legacy arguments (matrix_op-inl.h ReshapeParam.target_shape / keep_highest): a 0 is the one inferred dimension
        tgt = [int(v) for v in target_shape]
        if keep_highest:
            tgt[0] = data.shape[0]
        shape = tuple(-1 if v == 0 else v for v in tgt)
    if reverse in (True, 1, "1", "True", "true"):
        from .ndarray import _mx_reshape
        tgt = _mx_reshape(data.shape, tuple(shape), reverse=True)
        return _out(_op(lambda t: t.reshape(tgt), data), out)
    return _out(data.reshape(shape), out)


Reshape = reshape
_REG["Reshape"] = reshape


@register
def cast_storage(data, stype="default", out=None):
    """Storage-type conversion (tensor/cast_storage.cc); the gradient passes through unchanged (dense-backed)."""
    from . import sparse as _sp
    if data._t.device.type == "meta":
        return data                                        # shape inference: storage does not change the shape
    leaf = getattr(data, "_dense_leaf", None)
    if _ag.is_recording() and (leaf is not None or (getattr(data, "stype", "default") == "default" and data._t.requires_grad)):
        dense = _op(lambda t: t * 1, data)                 # taped identity on the dense view
        if stype == "default":
            return _out(dense, out)
        res = _sp.cast_storage(NDArray(dense._t.detach(), data.context), stype)
        res._dense_leaf = dense                            # operators downstream read the taped dense tensor
        return _out(res, out)
    return _out(_sp.cast_storage(data, stype), out)


@register(name="_sparse_retain")
def sparse_retain(data, indices, out=None):
    from . import sparse as _sp
    if data._t.device.type == "meta":
        return data
    leaf = getattr(data, "_dense_leaf", None)
    if _ag.is_recording() and leaf is not None:
        # taped form: rows that are not retained are multiplied by 0, the gradient reaches the retained rows only
        def g(t, i):
            keep = torch.zeros(t.shape[0], dtype=t.dtype, device=t.device)
            keep[i.long()] = 1
            return t * keep.reshape([-1] + [1] * (t.dim() - 1))
        dense = _op(g, data, indices)
        res = _sp.cast_storage(NDArray(dense._t.detach(), data.context), "row_sparse")
        res._dense_leaf = dense
        return _out(res, out)
    if not isinstance(data, _sp.RowSparseNDArray):
        data = _sp.cast_storage(data, "row_sparse")
    return _out(_sp.retain(data, indices), out)


@register
def reshape_like(lhs, rhs, lhs_begin=None, lhs_end=None, rhs_begin=None, rhs_end=None, out=None):
    """Reshape lhs like rhs; with the begin/end arguments only lhs.shape[lhs_begin:lhs_end] is replaced by
    rhs.shape[rhs_begin:rhs_end] (matrix_op.cc ReshapeLikeParam)."""
    if all(v is None for v in (lhs_begin, lhs_end, rhs_begin, rhs_end)):
        return _out(lhs.reshape(rhs.shape), out)
    ls, rs = list(lhs.shape), list(rhs.shape)
    lb, le, _ = np.s_[lhs_begin:lhs_end].indices(len(ls))
    rb, re_, _ = np.s_[rhs_begin:rhs_end].indices(len(rs))
    new = tuple(ls[:lb] + rs[rb:re_] + ls[le:])
    return _out(_op(lambda t: t.reshape(new), lhs), out)


@register
def flatten(data, out=None):
    return _out(_op(lambda t: t.reshape(t.shape[0], -1), data), out)


Flatten = flatten
_REG["Flatten"] = flatten


@register
def transpose(data, axes=None, out=None):
    return _out(data.transpose(*(axes or ())), out)


@register
def swapaxes(data, dim1=0, dim2=0, out=None):
    return _out(_op(lambda t: t.transpose(dim1, dim2), data), out)


SwapAxis = swapaxes
_REG["SwapAxis"] = swapaxes


@register
def expand_dims(data, axis, out=None):
    return _out(_op(lambda t: t.unsqueeze(axis), data), out)


@register
def squeeze(data, axis=None, out=None):
    return _out(data.squeeze(axis), out)


@register
def concat(*data, dim=1, out=None, **kw):
    dim = kw.get("axis", dim)
    return _out(_op(lambda *ts: torch.cat(ts, dim=dim), *data), out)


Concat = concat
_REG["Concat"] = concat


@register
def stack(*data, axis=0, out=None, **kw):
    return _out(_op(lambda *ts: torch.stack(ts, dim=axis), *data), out)


@register
def split(data, num_outputs, axis=1, squeeze_axis=False, out=None):
    def g(t):
        parts = torch.chunk(t, num_outputs, dim=axis)
        if squeeze_axis:
            parts = [p.squeeze(axis) for p in parts]
        return list(parts)
    r = _op(g, data)
    return r[0] if num_outputs == 1 else r


SliceChannel = split
_REG["SliceChannel"] = split


@register
def slice(data, begin, end, step=None, out=None):
    if isinstance(begin, (int, np.integer)) or begin is None:
        begin, end = (begin,), (end,)
    if isinstance(step, (int, np.integer)):
        step = (step,)
    if getattr(data, "stype", "default") == "csr" and (step is None or all(s in (None, 1) for s in step)):
        # row range of a csr matrix keeps the csr storage (matrix_op.cc SliceEx); column ranges go through the dense view
        from . import sparse as _sp
        b0, e0, _ = np.s_[begin[0]:end[0]].indices(data.shape[0])
        r = data[b0:e0]
        if len(begin) > 1 and (begin[1] not in (None, 0) or end[1] not in (None, data.shape[1])):
            b1, e1, _ = np.s_[begin[1]:end[1]].indices(data.shape[1])
            r = _sp.cast_storage(NDArray(r.todense()._t[:, b1:e1], data.context), "csr")
        return _out(r, out)

    def g(t):
        # numpy slice semantics incl. negative steps (matrix_op-inl.h SliceParam); torch cannot index with a negative
        # step, so those axes are gathered through an index vector
        r = t
        for i in range(len(begin)):
            st = step[i] if step is not None and i < len(step) and step[i] is not None else 1
            if st == 0:
                raise MXNetError("slice: step cannot be 0")
            b, e, st = np.s_[begin[i]:end[i]:st].indices(t.shape[i])
            if st > 0:
                r = r[(np.s_[:],) * i + (np.s_[b:e:st],)]
            else:
                idx = torch.arange(b, e, st, device=t.device)
                r = r.index_select(i, idx)
        return r
    return _out(_op(g, data), out)


crop = slice
_REG["crop"] = slice


@register
def slice_axis(data, axis, begin, end=None, out=None):
    return _out(data.slice_axis(axis, begin, end), out)


@register
def slice_like(data, shape_like, axes=None):
    def g(t, s):
        idx = [np.s_[:]] * t.dim()
        for a in (axes if axes else range(t.dim())):
            idx[a] = np.s_[:s.shape[a]]
        return t[tuple(idx)]
    return _op(g, data, shape_like)


@register
def tile(data, reps, out=None):
    reps = (reps,) if isinstance(reps, (int, np.integer)) else tuple(reps)
    if any(int(r) <= 0 for r in reps):
        raise MXNetError("tile: invalid reps value %s, every element must be positive" % (reps,))
    return _out(_op(lambda t: t.repeat(*reps) if len(reps) >= t.dim() else t.repeat(*([1] * (t.dim() - len(reps)) + list(reps))), data), out)


@register
def repeat(data, repeats, axis=None, out=None):
    return _out(_op(lambda t: torch.repeat_interleave(t if axis is not None else t.reshape(-1), repeats, dim=axis if axis is not None else 0), data), out)


@register
def flip(data, axis, out=None):
    return _out(_op(lambda t: torch.flip(t, [axis] if isinstance(axis, int) else list(axis)), data), out)


reverse = flip
_REG["reverse"] = flip


@register
def broadcast_to(data, shape, out=None):
    shape = tuple(s if s != 0 else d for s, d in zip(shape, data.shape))
    return _out(_op(lambda t: t.expand(*shape), data), out)


@register
def broadcast_like(lhs, rhs, lhs_axes=None, rhs_axes=None, out=None):
    """Broadcast lhs to the shape of rhs; with ``lhs_axes`` / ``rhs_axes`` only those axes of lhs take the sizes of the
    given axes of rhs (broadcast_reduce_op_value.cc BroadcastLikeParam)."""
    if lhs_axes is None and rhs_axes is None:
        return _out(broadcast_to(lhs, rhs.shape), out)
    la = [lhs_axes] if isinstance(lhs_axes, int) else list(lhs_axes)
    ra = [rhs_axes] if isinstance(rhs_axes, int) else list(rhs_axes)
    if len(la) != len(ra) or not la:
        raise MXNetError("broadcast_like: lhs_axes and rhs_axes must be non-empty and of equal length")
    shp = list(lhs.shape)
    for a, b in zip(la, ra):
        shp[a % lhs.ndim] = rhs.shape[b % rhs.ndim]
    return _out(broadcast_to(lhs, tuple(shp)), out)


@register
def broadcast_axis(data, axis, size, out=None):
    axs = [axis] if isinstance(axis, int) else list(axis)
    szs = [size] if isinstance(size, int) else list(size)
    shp = list(data.shape)
    for a, s in zip(axs, szs):
        shp[a] = s
    return _out(_op(lambda t: t.expand(*shp), data), out)


broadcast_axes = broadcast_axis
_REG["broadcast_axes"] = broadcast_axis


@register
def pad(data, mode="constant", pad_width=None, constant_value=0.0, out=None):
    def g(t):
        pw = list(pad_width)
        tp = []
        for i in reversed(range(len(pw) // 2)):
            tp += [pw[2 * i], pw[2 * i + 1]]
        m = {"constant": "constant", "edge": "replicate", "reflect": "reflect"}[mode]
        while len(tp) >= 2 and tp[-1] == 0 and tp[-2] == 0 and len(tp) > 2 * (t.dim() - 2):
            tp = tp[:-2]
        return F.pad(t, tp, mode=m, value=constant_value) if m == "constant" else F.pad(t, tp, mode=m)
    return _out(_op(g, data), out)


Pad = pad
_REG["Pad"] = pad


@register
def diag(data, k=0, axis1=0, axis2=1, out=None):
    """k-th diagonal of a matrix, or the matrix with a vector on its k-th diagonal (tensor/diag_op-inl.h); for more than
    two dimensions the diagonal of the (axis1, axis2) planes is appended as the last axis."""
    if data.ndim > 2 or (data.ndim == 2 and (axis1 % 2, axis2 % 2) == (1, 0)):
        return _out(_op(lambda t: torch.diagonal(t, k, axis1, axis2), data), out)
    if data.ndim == 2 and not (-data.shape[0] < k < data.shape[1]):
        raise MXNetError("diag: k=%d is outside of the %s matrix" % (k, str(data.shape)))
    if data.ndim == 0:
        raise MXNetError("diag: input must have at least 1 dimension")
    return _out(_op(lambda t: torch.diag(t, k) if t.dim() <= 2 else torch.diagonal(t, k), data), out)


@register
def shape_array(data):
    return NDArray(torch.tensor(data.shape, dtype=torch.int64, device=data._t.device), data._ctx)


@register
def size_array(data):
    return NDArray(torch.tensor([data.size], dtype=torch.int64, device=data._t.device), data._ctx)


# ---------------------------------------------------------------------------
# indexing
# ---------------------------------------------------------------------------

@register
def take(a, indices, axis=0, mode="clip", out=None):
    def g(t, i):
        i = i.long()
        ax = axis % t.dim()
        n = t.shape[ax]
        i = i.clamp(0, n - 1) if mode == "clip" else i % n
        return torch.index_select(t, ax, i.reshape(-1)).reshape(t.shape[:ax] + tuple(i.shape) + t.shape[ax + 1:])
    return _out(_op(g, a, indices), out)


@register
def pick(data, index, axis=-1, keepdims=False, mode="clip", out=None):
    def g(t, i):
        ax = axis % t.dim()
        i = i.long().clamp(0, t.shape[ax] - 1)
        r = torch.gather(t, ax, i.reshape(t.shape[:ax] + (1,) + t.shape[ax + 1:]))
        return r if keepdims else r.squeeze(ax)
    return _out(_op(g, data, index), out)


@register
def one_hot(indices, depth, on_value=1.0, off_value=0.0, dtype="float32", out=None):
    def g(i):
        if depth == 0:
            return torch.zeros(tuple(i.shape) + (0,), dtype=torch_dtype(dtype), device=i.device)
        r = F.one_hot(i.long().clamp(0, depth - 1), depth).to(torch_dtype(dtype))
        valid = ((i >= 0) & (i < depth)).unsqueeze(-1).to(r.dtype)
        r = r * valid
        if on_value != 1.0 or off_value != 0.0:
            r = r * (on_value - off_value) + off_value
        return r
    return _out(_op(g, indices), out)


@register
def gather_nd(data, indices):
    def g(t, i):
        i = i.long()
        return t[tuple(i[k] for k in range(i.shape[0]))]
    return _op(g, data, indices)


@register
def scatter_nd(data, indices, shape):
    def g(d, i):
        i = i.long()
        r = torch.zeros(tuple(shape), dtype=d.dtype, device=d.device)
        r[tuple(i[k] for k in range(i.shape[0]))] = d
        return r
    return _op(g, data, indices)


@register
def Embedding(data, weight, input_dim=None, output_dim=None, dtype="float32", sparse_grad=False, out=None):
    """src/operator/tensor/indexing_op.cu (Embedding / AddTakeGrad)."""
    return _out(_op(lambda i, w: F.embedding(i.long(), w), data, weight), out)


@register
def SequenceMask(data, sequence_length=None, use_sequence_length=False, value=0.0, axis=0, out=None):
    if not use_sequence_length or sequence_length is None:
        return data

    def g(t, l):
        T = t.shape[axis]
        ar = torch.arange(T, device=t.device)
        if axis == 0:
            m = ar[:, None] < l[None, :].to(ar.dtype)
        else:
            m = ar[None, :] < l[:, None].to(ar.dtype)
        m = m.reshape(m.shape + (1,) * (t.dim() - 2))
        return torch.where(m, t, torch.full_like(t, value))
    return _out(_op(g, data, sequence_length), out)


@register
def SequenceLast(data, sequence_length=None, use_sequence_length=False, axis=0):
    def g(t, l=None):
        if l is None:
            return t.select(axis, t.shape[axis] - 1)
        idx = (l.long() - 1)
        tt = t if axis == 0 else t.transpose(0, 1)
        return tt[idx, torch.arange(tt.shape[1], device=t.device)]
    if use_sequence_length and sequence_length is not None:
        return _op(g, data, sequence_length)
    return _op(g, data)


@register
def SequenceReverse(data, sequence_length=None, use_sequence_length=False, axis=0):
    def g(t, l=None):
        if l is None:
            return torch.flip(t, [0])
        T = t.shape[0]
        ar = torch.arange(T, device=t.device)[:, None]
        ll = l.long()[None, :]
        idx = torch.where(ar < ll, ll - 1 - ar, ar)
        return torch.gather(t, 0, idx.reshape(idx.shape + (1,) * (t.dim() - 2)).expand_as(t))
    if use_sequence_length and sequence_length is not None:
        return _op(g, data, sequence_length)
    return _op(g, data)


# ---------------------------------------------------------------------------
# linear algebra
# ---------------------------------------------------------------------------

@register
def dot(lhs, rhs, transpose_a=False, transpose_b=False, out=None, **kw):
    from . import sparse
    if isinstance(lhs, sparse.BaseSparseNDArray) or isinstance(rhs, sparse.BaseSparseNDArray):
        return _out(sparse.dot(lhs, rhs, transpose_a, transpose_b), out)

    def g(a, b):
        if transpose_a:
            a = a.t() if a.dim() == 2 else a.permute(*reversed(range(a.dim())))
        if transpose_b:
            b = b.t() if b.dim() == 2 else b.permute(*reversed(range(b.dim())))
        if a.dim() == 1 and b.dim() == 1:
            return torch.dot(a, b).reshape(1)
        return torch.tensordot(a, b, dims=1)
    return _out(_op(g, lhs, rhs), out)


@register
def batch_dot(lhs, rhs, transpose_a=False, transpose_b=False, out=None):
    """cublas gemmStridedBatched analogue (src/operator/linalg_impl.h:224-240)."""
    return _out(_op(lambda a, b: torch.matmul(a.transpose(-1, -2) if transpose_a else a,
                                              b.transpose(-1, -2) if transpose_b else b), lhs, rhs), out)


@register
def linalg_gemm2(A, B, transpose_a=False, transpose_b=False, alpha=1.0, axis=-2, out=None):
    """``axis``: which axis holds the matrix rows (la_op.cc: default -2; the columns are always the last axis)."""
    def g(a, b):
        a, b = _la_rows(a, axis), _la_rows(b, axis)
        r = alpha * torch.matmul(a.transpose(-1, -2) if transpose_a else a, b.transpose(-1, -2) if transpose_b else b)
        return _la_rows_back(r, axis)
    return _out(_op(g, A, B), out)


def _la_rows(t, axis):
    return t if axis in (-2, t.dim() - 2) else t.movedim(axis, -2)


def _la_rows_back(t, axis):
    return t if axis in (-2, t.dim() - 2) else t.movedim(-2, axis)


@register
def linalg_gemm(A, B, C, transpose_a=False, transpose_b=False, alpha=1.0, beta=1.0, axis=-2, out=None):
    def g(a, b, c):
        a, b, c = _la_rows(a, axis), _la_rows(b, axis), _la_rows(c, axis)
        r = alpha * torch.matmul(a.transpose(-1, -2) if transpose_a else a, b.transpose(-1, -2) if transpose_b else b) + beta * c
        return _la_rows_back(r, axis)
    return _out(_op(g, A, B, C), out)


@register
def linalg_potrf(A, out=None):
    return _out(_op(torch.linalg.cholesky, A), out)


@register
def linalg_trsm(A, B, transpose=False, rightside=False, alpha=1.0, out=None):
    def g(a, b):
        a_ = a.transpose(-1, -2) if transpose else a
        return alpha * torch.linalg.solve_triangular(a_, b, upper=transpose, left=not rightside)
    return _out(_op(g, A, B), out)


@register
def linalg_sumlogdiag(A, out=None):
    return _out(_op(lambda a: torch.log(torch.diagonal(a, dim1=-2, dim2=-1)).sum(-1), A), out)


@register
def linalg_syrk(A, transpose=False, alpha=1.0, out=None):
    return _out(_op(lambda a: alpha * (torch.matmul(a.transpose(-1, -2), a) if transpose else torch.matmul(a, a.transpose(-1, -2))), A), out)


@register(name="_contrib_div_sqrt_dim")
def div_sqrt_dim(data, out=None):
    """src/operator/contrib/transformer.cu:31-33."""
    return _out(_op(lambda t: t / math.sqrt(t.shape[-1]), data), out)


# ---------------------------------------------------------------------------
# neural-network ops
# ---------------------------------------------------------------------------

@register
def FullyConnected(data, weight, bias=None, num_hidden=None, no_bias=False, flatten=True, out=None):
    """Y = X.W^T + b (src/operator/nn/fully_connected-inl.h:72-180); tcgen05 GEMM on sm_100a."""
    from ..ops import functional as K

    def g(x, w, b=None):
        if flatten and x.dim() != 2:
            x = x.reshape(x.shape[0], -1)       # (N,) is N samples of one feature, (N, a, b) of a*b features
        if b is not None and b.dim() != 1:
            b = b.reshape(-1)                   # a (num_hidden, 1) bias is accepted like the reference's shape check does
        return K.linear(x, w, b)
    if bias is None or no_bias:
        return _out(_op(g, data, weight), out)
    return _out(_op(g, data, weight, bias), out)


@register
def Activation(data, act_type="relu", out=None):
    from ..ops.layernorm import bias_act

    def _native(t):   # one fused elementwise kernel forward, one backward (csrc/cuda/layernorm.cu)
        from ..ops.functional import act_any_layout
        return act_any_layout(t, lambda v: bias_act(v, None, act_type))
    if act_type == "softsign":
        return _out(_op(F.softsign, data), out)
    if act_type not in ("relu", "sigmoid", "tanh", "softrelu", "gelu"):
        raise MXNetError("unknown act_type " + str(act_type))
    return _out(_op(_native, data), out)


@register
def LeakyReLU(data, gamma=None, act_type="leaky", slope=0.25, lower_bound=0.125, upper_bound=0.334, out=None):
    def g(t, gm=None):
        if act_type in ("leaky", "elu", "selu", "gelu"):
            from ..ops.layernorm import bias_act
            return bias_act(t, None, act_type, slope)
        if act_type == "prelu":
            # leaky_relu-inl.h: a 1-D gamma is per channel (axis 1) or a single shared slope; a gamma with more
            # dimensions lines up with the LEADING dimensions of the data and broadcasts over the rest
            if gm.dim() <= 1:
                gb = gm.reshape([1, -1] + [1] * (t.dim() - 2)) if gm.numel() > 1 and t.dim() > 1 else gm.reshape([1] * t.dim() if gm.numel() == 1 else [-1])
            else:
                gb = gm.reshape(list(gm.shape) + [1] * (t.dim() - gm.dim()))
            return torch.where(t > 0, t, t * gb.to(t.dtype))
        if act_type == "rrelu":
            return F.rrelu(t, lower_bound, upper_bound, training=_ag.is_training())
        raise MXNetError("unknown act_type " + act_type)
    if gamma is not None:
        return _out(_op(g, data, gamma), out)
    return _out(_op(g, data), out)


def _tup(v, n):
    if v is None or v == ():
        return None
    if isinstance(v, (int, np.integer)):
        return (int(v),) * n
    return tuple(int(x) for x in v)


@register
def Convolution(data, weight, bias=None, kernel=None, stride=None, dilate=None, pad=None, num_filter=None,
                num_group=1, no_bias=False, layout=None, workspace=1024, cudnn_tune=None, cudnn_off=False, out=None):
    """src/operator/nn/convolution-inl.h:71-104.  NCHW logical layout; NHWC physical via channels_last."""
    nd_ = data.ndim - 2
    stride_, dil_, pad_ = _tup(stride, nd_) or (1,) * nd_, _tup(dilate, nd_) or (1,) * nd_, _tup(pad, nd_) or (0,) * nd_
    fn = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[nd_]

    def g(x, w, b=None):
        if nd_ == 2 and x.is_cuda:
            from ..ops import functional as K
            y = K.conv2d_nchw(x, w, b, stride_, pad_, dil_, num_group)     # tcgen05 GEMM for 1x1, channels_last cuDNN else
            if y is not None:
                return y
        return fn(x, w, b, stride_, pad_, dil_, num_group)
    if bias is None or no_bias:
        return _out(_op(g, data, weight), out)
    return _out(_op(g, data, weight, bias), out)


@register
def Deconvolution(data, weight, bias=None, kernel=None, stride=None, dilate=None, pad=None, adj=None, target_shape=None,
                  num_filter=None, num_group=1, no_bias=True, layout=None, workspace=512, out=None, **kw):
    nd_ = data.ndim - 2
    stride_, dil_, pad_ = _tup(stride, nd_) or (1,) * nd_, _tup(dilate, nd_) or (1,) * nd_, _tup(pad, nd_) or (0,) * nd_
    adj_ = _tup(adj, nd_) or (0,) * nd_
    fn = {1: F.conv_transpose1d, 2: F.conv_transpose2d, 3: F.conv_transpose3d}[nd_]

    def g(x, w, b=None):
        return fn(x, w, b, stride_, pad_, adj_, num_group, dil_)
    if bias is None or no_bias:
        return _out(_op(g, data, weight), out)
    return _out(_op(g, data, weight, bias), out)


@register
def Pooling(data, kernel=None, pool_type="max", global_pool=False, stride=None, pad=None,
            pooling_convention="valid", count_include_pad=True, cudnn_off=False, out=None, **kw):
    nd_ = data.ndim - 2

    def g(x):
        if global_pool:
            ax = tuple(range(2, x.dim()))
            if pool_type == "max":
                return x.amax(dim=ax, keepdim=True)
            if pool_type == "sum":
                return x.sum(dim=ax, keepdim=True)
            return x.mean(dim=ax, keepdim=True)
        k = _tup(kernel, nd_)
        s = _tup(stride, nd_) or (1,) * nd_
        p = (_tup(pad, nd_) or (0,) * nd_)[:nd_]
        ceil_mode = pooling_convention == "full"
        if pooling_convention == "same":
            # output = ceil(input / stride), padding implied on the right (pooling-inl.h: 1-D max pooling only, pad must be 0)
            if any(p) or pool_type != "max" or nd_ != 1:
                raise MXNetError("Pooling: the 'same' convention is for 1-D max pooling with zero pad")
            need = (-(-x.shape[2] // s[0]) - 1) * s[0] + k[0] - x.shape[2]
            if need > 0:
                x = F.pad(x, (0, need), value=float("-inf"))
            return F.max_pool1d(x, k, s, 0)
        if (pool_type == "max" and nd_ == 2 and x.is_cuda and not ceil_mode and k[0] == k[1] and k[0] in (2, 3)
                and s[0] == s[1] and p[0] == p[1] and p[0] < k[0] and x.dim() == 4 and not x.is_contiguous()
                and x.is_contiguous(memory_format=torch.channels_last)):
            from ..ops import functional as K
            if K._fastpath() and x.shape[1] % (16 // x.element_size()) == 0 and x.dtype in (torch.bfloat16, torch.float32):
                # layout chain already NHWC: window-index max pooling kernel (csrc/cuda/pool.cu), 1-byte argmax per output
                return K.max_pool_nhwc(x.permute(0, 2, 3, 1), k[0], s[0], p[0]).permute(0, 3, 1, 2)
        if pool_type == "max":
            return {1: F.max_pool1d, 2: F.max_pool2d, 3: F.max_pool3d}[nd_](x, k, s, p, ceil_mode=ceil_mode)
        fn = {1: F.avg_pool1d, 2: F.avg_pool2d, 3: F.avg_pool3d}[nd_]
        r = fn(x, k, s, p, ceil_mode=ceil_mode, count_include_pad=count_include_pad)
        if pool_type == "sum":
            r = r * float(np.prod(k))
        return r
    return _out(_op(g, data), out)


@register
def BatchNorm(data, gamma, beta, moving_mean, moving_var, eps=1e-3, momentum=0.9, fix_gamma=True,
              use_global_stats=False, output_mean_var=False, axis=1, cudnn_off=False, out=None, **kw):
    """src/operator/nn/batch_norm-inl.h:60-90.  Biased batch variance; running = running*momentum +
    batch*(1-momentum) (batch_norm.cu:359-362).  Moving stats are aux states, updated in place."""
    training = _ag.is_training() and not use_global_stats

    def g(x, gm, bt):
        mm, mv = moving_mean._t, moving_var._t
        if x.is_cuda and x.dim() == 4 and axis == 1 and not output_mean_var:
            from ..ops import functional as K
            y = K.batch_norm_nchw(x, gm, bt, mm, mv, eps, momentum, training, fix_gamma)   # fused statistics + apply kernels
            if y is not None:
                return y
        gm_ = torch.ones_like(gm) if fix_gamma else gm
        f64 = torch.float64
        if (x.dtype == f64) != (mm.dtype == f64) or (x.dtype == f64) != (gm_.dtype == f64):
            # float64 data with float32 parameters / statistics (or the reverse), which torch does not mix (the low-precision
            # data + fp32 statistics combination IS native and stays on the fast path): compute in the data's dtype and
            # write the statistics back
            mm_c, mv_c = mm.to(x.dtype), mv.to(x.dtype)
            r = F.batch_norm(x.movedim(axis, 1), mm_c, mv_c, gm_.to(x.dtype), bt.to(x.dtype), training, 1.0 - momentum, eps).movedim(1, axis)
            if training:
                with torch.no_grad():
                    mm.copy_(mm_c)
                    mv.copy_(mv_c)
            return r
        if axis != 1 and axis % x.dim() != 1:
            xt = x.movedim(axis, 1)
            r = F.batch_norm(xt, mm, mv, gm_, bt, training, 1.0 - momentum, eps)
            return r.movedim(1, axis)
        return F.batch_norm(x, mm, mv, gm_, bt, training, 1.0 - momentum, eps)
    return _out(_op(g, data, gamma, beta), out)


@register
def LayerNorm(data, gamma, beta, axis=-1, eps=1e-5, output_mean_var=False, out=None):
    """src/operator/nn/layer_norm.cu."""
    from ..ops import functional as K

    def g(x, gm, bt):
        if axis in (-1, x.dim() - 1):
            return K.layer_norm(x, gm, bt, eps)
        xt = x.movedim(axis, -1)
        return K.layer_norm(xt, gm, bt, eps).movedim(-1, axis)
    return _out(_op(g, data, gamma, beta), out)


@register
def InstanceNorm(data, gamma, beta, eps=1e-3, out=None):
    def f(x, g, b):
        red = tuple(range(2, x.dim()))
        xf = x.float()
        mean = xf.mean(dim=red, keepdim=True)
        var = xf.var(dim=red, unbiased=False, keepdim=True)
        shp = (1, -1) + (1,) * (x.dim() - 2)
        return ((xf - mean) * torch.rsqrt(var + eps) * g.float().reshape(shp) + b.float().reshape(shp)).to(x.dtype)
    return _out(_op(f, data, gamma, beta), out)


@register
def L2Normalization(data, eps=1e-10, mode="instance", out=None):
    def g(x):
        if mode == "instance":
            n = torch.sqrt((x * x).reshape(x.shape[0], -1).sum(1) + eps).reshape((-1,) + (1,) * (x.dim() - 1))
        elif mode == "channel":
            n = torch.sqrt((x * x).sum(1, keepdim=True) + eps)
        else:
            n = torch.sqrt((x * x).sum(tuple(range(2, x.dim())), keepdim=True) + eps)
        return x / n
    return _out(_op(g, data), out)


@register
def LRN(data, alpha=1e-4, beta=0.75, knorm=2.0, nsize=5, out=None):
    return _out(_op(lambda x: F.local_response_norm(x, nsize, alpha, beta, knorm), data), out)


@register
def Dropout(data, p=0.5, mode="training", axes=(), out=None, **kw):
    """Inverted dropout; ``axes``: the mask is shared (broadcast) along these axes -- variational dropout
    (src/operator/nn/dropout-inl.h:64-80)."""
    train = _ag.is_training() or mode == "always"
    axes_ = tuple(axes) if isinstance(axes, (list, tuple)) else ((axes,) if axes not in (None, ()) else ())

    def g(x):
        if not train or p == 0:
            return x
        if p >= 1:
            return x * float("nan")          # keep-probability 0: the reference's mask is 0 * (1/0) = NaN
        if not axes_:
            return F.dropout(x, p, training=True)
        shp = [1 if (i in axes_ or i - x.dim() in axes_) else s for i, s in enumerate(x.shape)]
        mask = (torch.rand(shp, device=x.device) >= p).to(x.dtype) / (1.0 - p)
        return x * mask
    return _out(_op(g, data), out)


@register
def softmax(data, axis=-1, temperature=None, out=None):
    return _out(_op(lambda t: torch.softmax(t / temperature if temperature else t, dim=axis), data), out)


@register
def log_softmax(data, axis=-1, temperature=None, out=None):
    return _out(_op(lambda t: torch.log_softmax(t / temperature if temperature else t, dim=axis), data), out)


@register
def SoftmaxActivation(data, mode="instance", out=None):
    return _out(_op(lambda t: torch.softmax(t, dim=1) if mode == "channel" else torch.softmax(t.reshape(t.shape[0], -1), 1).reshape(t.shape), data), out)


class _SoftmaxOutputFn(torch.autograd.Function):
    """SoftmaxOutput: fwd = softmax, bwd ignores the incoming gradient and writes
    (p - onehot)*grad_scale/norm (src/operator/softmax_output-inl.h:100-240)."""

    @staticmethod
    def forward(ctx, x, label, grad_scale, ignore_label, use_ignore, normalization, smooth_alpha, multi_output):
        from ..ops import functional as K
        if multi_output:
            p = torch.softmax(x, dim=1)
        else:
            p = K.softmax_rows(x.reshape(x.shape[0], -1)).reshape(x.shape)
        ctx.save_for_backward(p, label)
        ctx.cfg = (grad_scale, ignore_label, use_ignore, normalization, smooth_alpha, multi_output)
        return p

    @staticmethod
    def backward(ctx, gy):
        from ..ops import functional as K
        p, label = ctx.saved_tensors
        grad_scale, ignore_label, use_ignore, normalization, smooth_alpha, multi_output = ctx.cfg
        if multi_output:
            # (N, C, d1..) with label (N, d1..)
            C = p.shape[1]
            pm = p.movedim(1, -1).reshape(-1, C)
            lb = label.reshape(-1)
            g = K.softmax_output_grad(pm, lb, grad_scale, ignore_label, use_ignore, normalization, smooth_alpha)
            g = g.reshape(p.movedim(1, -1).shape).movedim(-1, 1)
        else:
            p2 = p.reshape(p.shape[0], -1)
            if label.shape == p.shape:   # dense (probability) label
                g = (p2 - label.reshape(p2.shape)) * grad_scale
                if normalization == "batch":
                    g = g / p2.shape[0]
            else:
                g = K.softmax_output_grad(p2, label.reshape(-1), grad_scale, ignore_label, use_ignore,
                                          normalization, smooth_alpha)
            g = g.reshape(p.shape)
        return g.to(p.dtype), None, None, None, None, None, None, None


@register
def SoftmaxOutput(data, label=None, grad_scale=1.0, ignore_label=-1.0, multi_output=False, use_ignore=False,
                  preserve_shape=False, normalization="null", out_grad=False, smooth_alpha=0.0, out=None, **kw):
    if label is None:
        return softmax(data, axis=1 if multi_output or not preserve_shape else -1, out=out)
    if preserve_shape and not multi_output:
        # softmax over the LAST axis, every leading position is its own row (softmax_output-inl.h preserve_shape)
        def gp(x, l):
            x2 = x.reshape(-1, x.shape[-1])
            l2 = l.reshape(x2.shape) if l.shape == x.shape else l.reshape(-1)
            return _SoftmaxOutputFn.apply(x2, l2, grad_scale, ignore_label, use_ignore, normalization, smooth_alpha, False).reshape(x.shape)
        return _out(_op(gp, data, label), out)
    return _out(_op(lambda x, l: _SoftmaxOutputFn.apply(x, l, grad_scale, ignore_label, use_ignore, normalization,
                                                        smooth_alpha, multi_output), data, label), out)


Softmax = SoftmaxOutput
_REG["Softmax"] = SoftmaxOutput


@register
def softmax_cross_entropy(data, label, out=None):
    """Sum over the batch of -log softmax(data)[label] (src/operator/loss_binary_op.cu)."""
    from ..ops import functional as K
    return _out(_op(lambda x, l: K.softmax_cross_entropy(x, l.long()).sum().reshape(1), data, label), out)


class _RegressionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, label, kind, grad_scale):
        y = torch.sigmoid(x) if kind == "logistic" else x.clone()
        ctx.save_for_backward(y, label)
        ctx.kind, ctx.gs = kind, grad_scale
        return y

    @staticmethod
    def backward(ctx, gy):
        y, label = ctx.saved_tensors
        d = y - label.reshape(y.shape).to(y.dtype)
        if ctx.kind == "mae":
            d = torch.sign(d)
        n = y.numel() // y.shape[0] if y.dim() > 0 else 1
        return d * (ctx.gs / n), None, None, None


@register
def LinearRegressionOutput(data, label, grad_scale=1.0, out=None):
    return _out(_op(lambda x, l: _RegressionFn.apply(x, l, "linear", grad_scale), data, label), out)


@register
def LogisticRegressionOutput(data, label, grad_scale=1.0, out=None):
    return _out(_op(lambda x, l: _RegressionFn.apply(x, l, "logistic", grad_scale), data, label), out)


@register
def MAERegressionOutput(data, label, grad_scale=1.0, out=None):
    return _out(_op(lambda x, l: _RegressionFn.apply(x, l, "mae", grad_scale), data, label), out)


class _MakeLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, grad_scale, normalization, valid_thresh):
        ctx.gs, ctx.norm = grad_scale, normalization
        ctx.save_for_backward(x)
        ctx.vt = valid_thresh
        return x.clone()

    @staticmethod
    def backward(ctx, gy):
        x, = ctx.saved_tensors
        g = torch.full_like(x, ctx.gs)
        if ctx.norm == "batch":
            g = g / x.shape[0]
        elif ctx.norm == "valid":
            g = g / torch.clamp((x > ctx.vt).sum(), min=1)
        return g, None, None, None


@register
def MakeLoss(data, grad_scale=1.0, valid_thresh=0.0, normalization="null", out=None):
    return _out(_op(lambda x: _MakeLossFn.apply(x, grad_scale, normalization, valid_thresh), data), out)


make_loss = MakeLoss
_REG["make_loss"] = MakeLoss


@register
def SVMOutput(data, label, margin=1.0, regularization_coefficient=1.0, use_linear=False, out=None):
    class _Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, l):
            ctx.save_for_backward(x, l)
            return x.clone()

        @staticmethod
        def backward(ctx, gy):
            x, l = ctx.saved_tensors
            y = F.one_hot(l.long(), x.shape[1]).to(x.dtype) * 2 - 1
            viol = (margin - y * x) > 0
            if use_linear:
                g = torch.where(viol, -y * regularization_coefficient, torch.zeros_like(x))
            else:
                g = torch.where(viol, -2 * (margin - y * x) * y * regularization_coefficient, torch.zeros_like(x))
            return g, None
    return _out(_op(lambda x, l: _Fn.apply(x, l), data, label), out)


@register
def UpSampling(*data, scale=2, sample_type="nearest", num_args=1, num_filter=0, multi_input_mode="concat", workspace=512, out=None, **kw):
    """nn/upsampling-inl.h: 'nearest' brings every input to scale x the FIRST input's size (input k by its own integer
    factor) and concatenates them on the channel axis (or sums them, ``multi_input_mode='sum'``); 'bilinear' is a grouped
    transposed convolution of ``data`` with the ``weight`` input (kernel 2*scale - scale%2, stride scale, pad
    ceil((scale-1)/2), one group per channel)."""
    scale = int(scale)
    if sample_type == "nearest":
        def g(x, *rest):
            oh = x.shape[2] * scale
            ups = [F.interpolate(t, scale_factor=oh // t.shape[2], mode="nearest") if oh != t.shape[2] else t for t in (x,) + rest]
            if len(ups) == 1:
                return ups[0]
            if multi_input_mode == "sum":
                r = ups[0]
                for u in ups[1:]:
                    r = r + u
                return r
            return torch.cat(ups, 1)
        return _out(_op(g, *data), out)
    if len(data) < 2:
        return _out(_op(lambda x: F.interpolate(x, scale_factor=scale, mode="bilinear"), data[0]), out)
    k, pd = 2 * scale - scale % 2, int(math.ceil((scale - 1) / 2.0))
    return _out(_op(lambda x, w: F.conv_transpose2d(x, w.to(x.dtype), None, stride=scale, padding=pd, groups=x.shape[1]), data[0], data[1]), out)


@register(name="_contrib_BilinearResize2D")
def BilinearResize2D(data, height, width, out=None):
    return _out(_op(lambda x: F.interpolate(x, size=(height, width), mode="bilinear", align_corners=True), data), out)


@register(name="_contrib_AdaptiveAvgPooling2D")
def AdaptiveAvgPooling2D(data, output_size=1, out=None):
    return _out(_op(lambda x: F.adaptive_avg_pool2d(x, output_size), data), out)


@register
def GridGenerator(data, transform_type="affine", target_shape=None):
    def g(t):
        if transform_type == "affine":
            return F.affine_grid(t.reshape(-1, 2, 3), (t.shape[0], 1) + tuple(target_shape), align_corners=True).permute(0, 3, 1, 2)
        H, W = t.shape[2], t.shape[3]
        ys, xs = torch.meshgrid(torch.linspace(-1, 1, H, device=t.device), torch.linspace(-1, 1, W, device=t.device), indexing="ij")
        base = torch.stack([xs, ys])[None]
        return base + t / torch.tensor([(W - 1) / 2.0, (H - 1) / 2.0], device=t.device).reshape(1, 2, 1, 1)
    return _op(g, data)


@register
def BilinearSampler(data, grid, out=None):
    return _out(_op(lambda x, g: F.grid_sample(x, g.permute(0, 2, 3, 1), mode="bilinear", padding_mode="zeros", align_corners=True), data, grid), out)


@register
def SpatialTransformer(data, loc, target_shape=None, transform_type="affine", sampler_type="bilinear", out=None):
    g = GridGenerator(loc, "affine", target_shape)
    return BilinearSampler(data, g, out=out)


def _clip(v, lo, hi):
    return lo if v < lo else (hi if v > hi else v)        # (the module-level min / max are the reduction operators)


@register
def ROIPooling(data, rois, pooled_size, spatial_scale, out=None):
    """Max pooling of every ROI to a fixed (ph, pw) grid (src/operator/roi_pooling.cc): ROI corners are rounded to the
    feature map, a ROI spans end - start + 1 cells (at least 1), bin (i, j) covers [floor(i*bh), ceil((i+1)*bh)) clipped to
    the map; an empty bin yields 0."""
    ph, pw = (int(pooled_size),) * 2 if isinstance(pooled_size, (int, np.integer)) else (int(pooled_size[0]), int(pooled_size[1]))

    def g(x, r):
        R, (C, H, W) = r.shape[0], x.shape[1:]
        if x.device.type == "meta":
            return x.new_empty((R, C, ph, pw))
        rr = r.detach().double().cpu()
        outs = []
        for k in range(R):
            b = int(rr[k, 0])
            x1, y1, x2, y2 = [int(math.floor(float(v) * float(spatial_scale) + 0.5)) for v in rr[k, 1:5]]
            rh, rw = (y2 - y1 + 1 if y2 - y1 + 1 > 1 else 1), (x2 - x1 + 1 if x2 - x1 + 1 > 1 else 1)
            bh, bw = rh / float(ph), rw / float(pw)
            rows = []
            for i in range(ph):
                hs, he = int(math.floor(i * bh)) + y1, int(math.ceil((i + 1) * bh)) + y1
                hs, he = _clip(hs, 0, H), _clip(he, 0, H)
                cols = []
                for j in range(pw):
                    ws, we = int(math.floor(j * bw)) + x1, int(math.ceil((j + 1) * bw)) + x1
                    ws, we = _clip(ws, 0, W), _clip(we, 0, W)
                    if he <= hs or we <= ws:
                        cols.append(x.new_zeros((C,)))
                    else:
                        cols.append(x[b, :, hs:he, ws:we].amax(dim=(1, 2)))
                rows.append(torch.stack(cols, -1))
            outs.append(torch.stack(rows, -2))
        return torch.stack(outs, 0) if outs else x.new_zeros((0, C, ph, pw))
    return _out(_op(g, data, rois), out)


@register(name="_contrib_ROIAlign")
def ROIAlign(data, rois, pooled_size, spatial_scale=1.0, sample_ratio=0, out=None, **kw):
    """ROI align (src/operator/contrib/roi_align.cc): every output bin averages a grid of bilinear samples
    (``sample_ratio`` per side, or ceil(roi extent / pooled extent) when <= 0); samples further than one pixel outside the
    map contribute 0, coordinates are clamped at the border.  No gradient reaches ``rois``.  One gather per corner over the
    whole sample grid of a ROI -- autograd turns it into the scatter-add of the backward pass."""
    ph, pw = (int(pooled_size),) * 2 if isinstance(pooled_size, (int, np.integer)) else (int(pooled_size[0]), int(pooled_size[1]))

    def g(x, r):
        R, (C, H, W) = r.shape[0], x.shape[1:]
        if x.device.type == "meta":
            return x.new_empty((R, C, ph, pw))
        rr = r.detach().double().cpu()
        outs = []
        for k in range(R):
            b = int(rr[k, 0])
            sw, sh, ew, eh = [float(v) * float(spatial_scale) for v in rr[k, 1:5]]
            roi_w, roi_h = ew - sw if ew - sw > 1.0 else 1.0, eh - sh if eh - sh > 1.0 else 1.0
            bin_h, bin_w = roi_h / ph, roi_w / pw
            gh = int(sample_ratio) if int(sample_ratio) > 0 else int(math.ceil(roi_h / ph))
            gw = int(sample_ratio) if int(sample_ratio) > 0 else int(math.ceil(roi_w / pw))
            ys = sh + (torch.arange(ph, dtype=torch.float64)[:, None] * bin_h + (torch.arange(gh, dtype=torch.float64)[None, :] + 0.5) * bin_h / gh).reshape(-1)
            xs = sw + (torch.arange(pw, dtype=torch.float64)[:, None] * bin_w + (torch.arange(gw, dtype=torch.float64)[None, :] + 0.5) * bin_w / gw).reshape(-1)

            def corners(c, n):
                dead = (c < -1.0) | (c > n)
                c = c.clamp_min(0.0)
                lo = c.floor().long()
                top = lo >= n - 1
                lo = torch.where(top, torch.full_like(lo, n - 1), lo)
                hi = torch.where(top, lo, lo + 1)
                frac = torch.where(top, torch.zeros_like(c), c - lo.to(c.dtype))
                return lo, hi, frac, dead
            yl, yh, ly, ydead = corners(ys, H)
            xl, xh, lx, xdead = corners(xs, W)
            dev = x.device
            wy = torch.stack([1.0 - ly, ly]).to(dev, x.dtype)                    # (2, ph*gh)
            wx = torch.stack([1.0 - lx, lx]).to(dev, x.dtype)
            live = ((~ydead)[:, None] & (~xdead)[None, :]).to(dev, x.dtype)      # (ph*gh, pw*gw)
            img = x[b]
            acc = 0
            for iy, yy in enumerate((yl, yh)):
                rows = img.index_select(1, yy.to(dev))
                for ix, xx in enumerate((xl, xh)):
                    acc = acc + rows.index_select(2, xx.to(dev)) * (wy[iy][:, None] * wx[ix][None, :])
            acc = acc * live
            outs.append(acc.reshape(C, ph, gh, pw, gw).mean(dim=(2, 4)))
        return torch.stack(outs, 0) if outs else x.new_zeros((0, C, ph, pw))
    return _out(_op(g, data, rois), out)


@register
def CTCLoss(data, label, data_lengths=None, label_lengths=None, use_data_lengths=False, use_label_lengths=False,
            blank_label="first", out=None):
    use_dl = bool(use_data_lengths) and data_lengths is not None
    use_ll = bool(use_label_lengths) and label_lengths is not None

    def g(x, l, *lens):
        T, N, C = x.shape
        if x.device.type == "meta":                      # shape inference
            return x.new_empty((N,))
        lens = list(lens)
        dl = lens.pop(0) if use_dl else None
        ll = lens.pop(0) if use_ll else None
        if dl is not None:
            # frames past a sequence's length take no part (they may hold anything, NaN included): zero value, zero gradient
            live = (torch.arange(T, device=x.device)[:, None] < dl.long()[None, :])[:, :, None]
            x = torch.where(live, x, torch.zeros_like(x))
        lp = torch.log_softmax(x.float(), dim=2)
        blank = 0 if blank_label == "first" else C - 1
        li = l.long()
        if ll is None:
            # ctc_loss-inl.h LabelTensorToPackedVector: a label sequence ends at the FIRST padding value (0 when the blank
            # is the first class, -1 when it is the last)
            pad = 0 if blank_label == "first" else -1
            is_pad = (li == pad)
            first = torch.where(is_pad.any(1), is_pad.float().argmax(1), torch.full((N,), li.shape[1], device=li.device))
            ll_ = first.long()
        else:
            ll_ = ll.long()
        dl_ = dl.long() if dl is not None else torch.full((N,), T, dtype=torch.long, device=x.device)
        li = li.clamp_min(0)
        return F.ctc_loss(lp, li, dl_, ll_, blank=blank, reduction="none", zero_infinity=True).to(x.dtype)
    args = [data, label]
    if use_dl:
        args.append(data_lengths)
    if use_ll:
        args.append(label_lengths)
    return _out(_op(lambda *a: g(*a), *args), out)


ctc_loss = CTCLoss
_REG["ctc_loss"] = CTCLoss
_REG["_contrib_ctc_loss"] = CTCLoss
_REG["_contrib_CTCLoss"] = CTCLoss


@register
def RNN(data, parameters, state, state_cell=None, state_size=None, num_layers=1, bidirectional=False, mode="lstm",
        p=0.0, state_outputs=False, out=None, **kw):
    """Fused RNN (src/operator/rnn.cu -> cuDNN RNN); parameters is the flat cuDNN-layout vector."""
    from ..gluon.rnn import _fused_rnn_forward
    return _fused_rnn_forward(data, parameters, state, state_cell, state_size, num_layers, bidirectional, mode, p,
                              state_outputs)


# ---------------------------------------------------------------------------
# random
# ---------------------------------------------------------------------------

@register
def shuffle(data, out=None):
    return _out(_op(lambda t: t[torch.randperm(t.shape[0], device=t.device)], data), out)


# ---------------------------------------------------------------------------
# optimizer update ops (src/operator/optimizer_op-inl.h).  ``out`` is normally the weight itself.
# All hyper-parameters are plain kernel arguments: lr changes never recompile.
# ---------------------------------------------------------------------------

def _prep_grad(g, rescale_grad, clip_gradient):
    g = g.float() * rescale_grad
    if clip_gradient is not None and clip_gradient >= 0:
        g = torch.clamp(g, -clip_gradient, clip_gradient)
    return g


def _fused_ok(*ts):
    return all(t.is_cuda and t.is_contiguous() for t in ts)


@register
def sgd_update(weight, grad, lr, wd=0.0, rescale_grad=1.0, clip_gradient=-1.0, lazy_update=True, out=None):
    """w = (1-lr*wd)*w - lr*clip(rescale*g)   (optimizer_op-inl.h:86-125)"""
    out = out if out is not None else weight
    w, g = weight._t, grad._t
    with torch.no_grad():
        if _fused_ok(w, g) and out is weight:
            from ..ops import optim as KO
            KO.sgd_update(w, g, None, None, lr, wd, 0.0, rescale_grad, clip_gradient)
        else:
            gg = _prep_grad(g, rescale_grad, clip_gradient)
            out._t.copy_(((1.0 - lr * wd) * w.float() - lr * gg).to(w.dtype))
    return out


@register
def sgd_mom_update(weight, grad, mom, lr, momentum=0.0, wd=0.0, rescale_grad=1.0, clip_gradient=-1.0,
                   lazy_update=True, out=None):
    """m = mu*m - lr*wd*w - lr*clip(rescale*g); w += m   (optimizer_op-inl.h:305-353)"""
    out = out if out is not None else weight
    w, g, m = weight._t, grad._t, mom._t
    with torch.no_grad():
        if _fused_ok(w, g, m) and out is weight:
            from ..ops import optim as KO
            KO.sgd_update(w, g, m, None, lr, wd, momentum, rescale_grad, clip_gradient)
        else:
            gg = _prep_grad(g, rescale_grad, clip_gradient)
            m.copy_((momentum * m.float() - lr * wd * w.float() - lr * gg).to(m.dtype))
            out._t.copy_((w.float() + m.float()).to(w.dtype))
    return out


@register
def mp_sgd_update(weight, grad, weight32, lr, wd=0.0, rescale_grad=1.0, clip_gradient=-1.0, lazy_update=True, out=None):
    """Low-precision weight/grad + fp32 master (optimizer_op-inl.h:359-400)."""
    out = out if out is not None else weight
    w, g, w32 = weight._t, grad._t, weight32._t
    with torch.no_grad():
        if _fused_ok(w, g, w32) and out is weight:
            from ..ops import optim as KO
            KO.sgd_update(w, g, None, w32, lr, wd, 0.0, rescale_grad, clip_gradient)
        else:
            gg = _prep_grad(g, rescale_grad, clip_gradient)
            w32.copy_((1.0 - lr * wd) * w32 - lr * gg)
            out._t.copy_(w32.to(w.dtype))
    return out


@register
def mp_sgd_mom_update(weight, grad, mom, weight32, lr, momentum=0.0, wd=0.0, rescale_grad=1.0, clip_gradient=-1.0,
                      lazy_update=True, out=None):
    """fp32 momentum + fp32 master, low-precision weight written back (optimizer_op-inl.h:403-448)."""
    out = out if out is not None else weight
    w, g, m, w32 = weight._t, grad._t, mom._t, weight32._t
    with torch.no_grad():
        if _fused_ok(w, g, m, w32) and out is weight:
            from ..ops import optim as KO
            KO.sgd_update(w, g, m, w32, lr, wd, momentum, rescale_grad, clip_gradient)
        else:
            gg = _prep_grad(g, rescale_grad, clip_gradient)
            m.copy_(momentum * m - lr * wd * w32 - lr * gg)
            w32.add_(m)
            out._t.copy_(w32.to(w.dtype))
    return out


@register
def adam_update(weight, grad, mean, var, lr, beta1=0.9, beta2=0.999, epsilon=1e-8, wd=0.0, rescale_grad=1.0,
                clip_gradient=-1.0, lazy_update=True, out=None):
    """g = rescale*g + wd*w; m,v EMA; w -= lr*m/(sqrt(v)+eps)   (optimizer_op-inl.h:841-876)"""
    out = out if out is not None else weight
    w, g, m, v = weight._t, grad._t, mean._t, var._t
    with torch.no_grad():
        if _fused_ok(w, g, m, v) and out is weight and m.dtype is torch.float32:
            from ..ops import optim as KO
            KO.adam_update(w, g, m, v, None, lr, beta1, beta2, epsilon, wd, rescale_grad, clip_gradient)
        else:
            gg = g.float() * rescale_grad + wd * w.float()
            if clip_gradient is not None and clip_gradient >= 0:
                gg = torch.clamp(gg, -clip_gradient, clip_gradient)
            m.copy_((beta1 * m.float() + (1 - beta1) * gg).to(m.dtype))
            v.copy_((beta2 * v.float() + (1 - beta2) * gg * gg).to(v.dtype))
            out._t.copy_((w.float() - lr * m.float() / (torch.sqrt(v.float()) + epsilon)).to(w.dtype))
    return out


@register
def signsgd_update(weight, grad, lr, wd=0.0, rescale_grad=1.0, clip_gradient=-1.0, out=None):
    out = out if out is not None else weight
    with torch.no_grad():
        w = weight._t.float()
        out._t.copy_(((1 - lr * wd) * w - lr * torch.sign(grad._t.float())).to(weight._t.dtype))
    return out


@register
def signum_update(weight, grad, mom, lr, momentum=0.0, wd=0.0, rescale_grad=1.0, clip_gradient=-1.0, wd_lh=0.0, out=None):
    out = out if out is not None else weight
    with torch.no_grad():
        w = weight._t.float()
        gg = _prep_grad(grad._t, rescale_grad, clip_gradient)
        mom._t.copy_(momentum * mom._t - (1 - momentum) * wd * w - (1 - momentum) * gg)
        out._t.copy_(((1 - lr * wd_lh) * w + lr * torch.sign(mom._t)).to(weight._t.dtype))
    return out


@register
def rmsprop_update(weight, grad, n, lr, gamma1=0.95, epsilon=1e-8, wd=0.0, rescale_grad=1.0, clip_gradient=-1.0,
                   clip_weights=-1.0, out=None):
    out = out if out is not None else weight
    with torch.no_grad():
        w = weight._t.float()
        gg = grad._t.float() * rescale_grad + wd * w
        if clip_gradient is not None and clip_gradient >= 0:
            gg = torch.clamp(gg, -clip_gradient, clip_gradient)
        n._t.copy_((1 - gamma1) * gg * gg + gamma1 * n._t)
        w = w - lr * gg / torch.sqrt(n._t + epsilon)
        if clip_weights is not None and clip_weights >= 0:
            w = torch.clamp(w, -clip_weights, clip_weights)
        out._t.copy_(w.to(weight._t.dtype))
    return out


@register
def rmspropalex_update(weight, grad, n, g, delta, lr, gamma1=0.95, gamma2=0.9, epsilon=1e-8, wd=0.0, rescale_grad=1.0,
                       clip_gradient=-1.0, clip_weights=-1.0, out=None):
    out = out if out is not None else weight
    with torch.no_grad():
        w = weight._t.float()
        gg = grad._t.float() * rescale_grad + wd * w
        if clip_gradient is not None and clip_gradient >= 0:
            gg = torch.clamp(gg, -clip_gradient, clip_gradient)
        n._t.copy_((1 - gamma1) * gg * gg + gamma1 * n._t)
        g._t.copy_((1 - gamma1) * gg + gamma1 * g._t)
        delta._t.copy_(gamma2 * delta._t - lr * gg / torch.sqrt(n._t - g._t * g._t + epsilon))
        w = w + delta._t
        if clip_weights is not None and clip_weights >= 0:
            w = torch.clamp(w, -clip_weights, clip_weights)
        out._t.copy_(w.to(weight._t.dtype))
    return out


@register
def ftrl_update(weight, grad, z, n, lr, lamda1=0.01, beta=1.0, wd=0.0, rescale_grad=1.0, clip_gradient=-1.0, out=None):
    out = out if out is not None else weight
    with torch.no_grad():
        w = weight._t.float()
        gg = _prep_grad(grad._t, rescale_grad, clip_gradient)
        z._t.add_(gg - (torch.sqrt(n._t + gg * gg) - torch.sqrt(n._t)) * w / lr)
        n._t.add_(gg * gg)
        w = (torch.sign(z._t) * lamda1 - z._t) / ((beta + torch.sqrt(n._t)) / lr + wd) * (z._t.abs() > lamda1).float()
        out._t.copy_(w.to(weight._t.dtype))
    return out


@register
def ftml_update(weight, grad, d, v, z, lr, beta1=0.6, beta2=0.999, epsilon=1e-8, t=1, wd=0.0, rescale_grad=1.0,
                clip_grad=-1.0, out=None):
    out = out if out is not None else weight
    with torch.no_grad():
        w = weight._t.float()
        gg = grad._t.float() * rescale_grad + wd * w
        if clip_grad is not None and clip_grad >= 0:
            gg = torch.clamp(gg, -clip_grad, clip_grad)
        v._t.copy_(beta2 * v._t + (1 - beta2) * gg * gg)
        d_t = (1 - beta1 ** t) / lr * (torch.sqrt(v._t / (1 - beta2 ** t)) + epsilon)
        sigma = d_t - beta1 * d._t
        z._t.copy_(beta1 * z._t + (1 - beta1) * gg - sigma * w)
        d._t.copy_(d_t)
        out._t.copy_((-z._t / d_t).to(weight._t.dtype))
    return out


@register(name="_contrib_group_adagrad_update")
def group_adagrad_update(weight, grad, history, lr, rescale_grad=1.0, clip_gradient=-1.0, epsilon=1e-5, out=None):
    out = out if out is not None else weight
    with torch.no_grad():
        gg = _prep_grad(grad._t, rescale_grad, clip_gradient)
        history._t.add_((gg * gg).mean(dim=tuple(range(1, gg.dim())), keepdim=True).reshape(history._t.shape))
        out._t.copy_((weight._t.float() - lr * gg / torch.sqrt(history._t.reshape((-1,) + (1,) * (gg.dim() - 1)) + epsilon)).to(weight._t.dtype))
    return out


def get(name):
    return _REG.get(name)


# ----------------------------------------------------------------------------- row_sparse gradients / weights in update ops
def _sparse_aware_update(name, n_states, lazy_capable):
    """Wrap a dense update operator so it accepts row_sparse gradients / weights / states like the reference kernels
    (optimizer_op-inl.h: SGDDnsRspKernel, SGDMomStdDnsRspDnsKernel, AdamStdDnsRspDnsKernel, FtrlDnsRspDnsKernel ...):

    * ``lazy_update`` (default True) with a row_sparse gradient touches ONLY the rows present in the gradient -- weight
      decay and momentum decay of absent rows are skipped;
    * otherwise ("standard" update) absent rows see a zero gradient;
    * row_sparse weights / states are updated through a dense round trip and written back in place."""
    from . import sparse as _spm
    dense_fn = _REG[name]

    def update(weight, grad, *args, **kwargs):
        is_sp = lambda a: isinstance(a, _spm.BaseSparseNDArray)   # noqa: E731
        states = list(args[:n_states])
        rest = args[n_states:]
        if not (is_sp(weight) or is_sp(grad) or any(is_sp(s) for s in states)):
            return dense_fn(weight, grad, *args, **kwargs)
        out = kwargs.pop("out", None)
        lazy = kwargs.get("lazy_update", True) if lazy_capable == "flag" else (lazy_capable == "always")
        dw = weight.todense() if is_sp(weight) else weight
        ds = [s.todense() if is_sp(s) else s for s in states]
        if isinstance(grad, _spm.RowSparseNDArray) and lazy:
            rows = grad._indices
            if rows.numel():
                w_rows = NDArray(dw._t.index_select(0, rows), dw.context)
                s_rows = [NDArray(s._t.index_select(0, rows), s.context) for s in ds]
                dense_fn(w_rows, NDArray(grad._data, grad.context), *s_rows, *rest, **kwargs)
                with torch.no_grad():
                    dw._t.index_copy_(0, rows, w_rows._t)
                    for s, sr in zip(ds, s_rows):
                        s._t.index_copy_(0, rows, sr._t)
        else:
            dg = grad.todense() if is_sp(grad) else grad
            dense_fn(dw, dg, *ds, *rest, **kwargs)
        for orig, dense in zip([weight] + states, [dw] + ds):
            if is_sp(orig):
                orig[:] = dense
        if out is not None and out is not weight:
            out[:] = dw
            return out
        return weight
    update.__name__ = name
    update.__doc__ = dense_fn.__doc__
    return update


for _name, _ns, _lazy in (("sgd_update", 0, "flag"), ("sgd_mom_update", 1, "flag"), ("mp_sgd_update", 1, "flag"),
                          ("mp_sgd_mom_update", 2, "flag"), ("adam_update", 2, "flag"), ("ftrl_update", 2, "always"),
                          ("signsgd_update", 0, "never"), ("signum_update", 1, "never"), ("rmsprop_update", 1, "never"),
                          ("rmspropalex_update", 3, "never"), ("ftml_update", 3, "never"),
                          ("_contrib_group_adagrad_update", 1, "always")):
    if _name in _REG:
        _REG[_name] = _sparse_aware_update(_name, _ns, _lazy)
        if _name in globals():
            globals()[_name] = _REG[_name]


# ----------------------------------------------------------------------------- pre-NNVM helpers still exported by mx.nd
@register
def onehot_encode(indices, out):
    """out[i, indices[i]] = 1, everything else 0 (legacy ndarray function, ndarray.py:onehot_encode)."""
    with torch.no_grad():
        o = out._t
        o.zero_()
        o.scatter_(1, indices._t.long().reshape(-1, 1), 1.0)
    return out


@register
def choose_element_0index(lhs, rhs, out=None):
    """out[i] = lhs[i, rhs[i]]."""
    return _out(_op(lambda a, idx: a.gather(1, idx.long().reshape(-1, 1)).reshape(-1), lhs, rhs), out)


@register
def fill_element_0index(lhs, mhs, rhs, out=None):
    """Copy of lhs with out[i, rhs[i]] = mhs[i]."""
    def g(a, v, idx):
        r = a.clone()
        r.scatter_(1, idx.long().reshape(-1, 1), v.reshape(-1, 1).to(a.dtype))
        return r
    return _out(_op(g, lhs, mhs, rhs), out)


class CachedOp(object):
    """Callable compiled from a Symbol (ndarray CachedOp of the reference, src/imperative/cached_op.cc): inputs are bound
    positionally in ``sym.list_inputs()`` order.  Here a thin wrapper over the graph executor's evaluation loop, so the
    BatchNorm(+add)+ReLU grouping of the executor applies to hybridized Gluon blocks that go through it."""

    def __init__(self, sym, flags=()):
        self._sym = sym
        self._names = sym.list_inputs()
        self._flags = dict(flags)

    def __call__(self, *args, **kwargs):
        out = kwargs.pop("out", None)
        if kwargs:
            raise TypeError("CachedOp.__call__ got unexpected keyword argument(s): %s" % ", ".join(kwargs))
        if len(args) != len(self._names):
            raise MXNetError("CachedOp expects %d inputs (%s), got %d" % (len(self._names), self._names, len(args)))
        res = self._sym.eval_nd(dict(zip(self._names, args)))
        if out is not None:
            outs = out if isinstance(out, (list, tuple)) else [out]
            for o, r in zip(outs, res):
                o[:] = r
            return out
        return res[0] if len(res) == 1 else res
