"""NDArray: the framework's array handle, a thin view over ``torch.Tensor``.

Reference surface: python/mxnet/ndarray/ndarray.py (NDArray methods),
include/mxnet/ndarray.h:82-1427 (chunk + engine var + lazy alloc).  Design
difference: torch owns storage and stream ordering, so an NDArray is just
``(tensor, ctx)``; the "engine variable" of the reference collapses into the
CUDA stream the tensor was produced on, and ``wait_to_read`` is a stream/event
wait (see mxb200/engine.py for the explicit dependency engine used by the
kvstore and IO pipelines).
"""
import numpy as np
import torch

from ..base import MXNetError, np_dtype, torch_dtype, numeric_types, bfloat16
from ..context import Context, current_context, from_torch_device
from .. import autograd as _ag

__all__ = ["NDArray", "array", "zeros", "ones", "empty", "full", "arange", "zeros_like",
           "ones_like", "from_torch", "waitall", "concatenate", "moveaxis"]


def _ctx_of(ctx):
    if ctx is None:
        return current_context()
    if isinstance(ctx, Context):
        return ctx
    if isinstance(ctx, (torch.device, str)):
        return from_torch_device(ctx)
    raise TypeError("bad context %r" % (ctx,))


def _unwrap(x):
    return x._t if isinstance(x, NDArray) else x


def _result_ctx(*args):
    for a in args:
        if isinstance(a, NDArray):
            return a._ctx
    return current_context()


def _pin_if(t, ctx):
    if ctx.device_typeid == 3 and torch.cuda.is_available() and not t.is_pinned():
        return t.pin_memory()
    return t


# storage-type / dtype / grad_req code tables of the reference's C ABI (ndarray.py:52-98); kept because user code and
# serialised files refer to the integer codes
_STORAGE_TYPE_UNDEFINED, _STORAGE_TYPE_DEFAULT, _STORAGE_TYPE_ROW_SPARSE, _STORAGE_TYPE_CSR = -1, 0, 1, 2
_STORAGE_TYPE_STR_TO_ID = {"undefined": -1, "default": 0, "row_sparse": 1, "csr": 2}
_STORAGE_TYPE_ID_TO_STR = {v: k for k, v in _STORAGE_TYPE_STR_TO_ID.items()}
_DTYPE_NP_TO_MX = {None: -1, np.float32: 0, np.float64: 1, np.float16: 2, np.uint8: 3, np.int32: 4, np.int8: 5, np.int64: 6}
_DTYPE_MX_TO_NP = {-1: None, 0: np.float32, 1: np.float64, 2: np.float16, 3: np.uint8, 4: np.int32, 5: np.int8, 6: np.int64}
_GRAD_REQ_MAP = {"null": 0, "write": 1, "add": 3}


class NDArray(object):
    """n-dimensional array on a Context.  ``arr.data_`` / ``arr._t`` is the torch tensor."""
    __slots__ = ("_t", "_ctx", "_grad_req", "_grad_hook", "_bw_epoch", "_grad_stype", "__weakref__")
    __array_priority__ = 1000.0

    def __init__(self, t, ctx=None):
        if isinstance(t, NDArray):
            t, ctx = t._t, ctx or t._ctx
        self._t = t
        self._ctx = ctx if ctx is not None else from_torch_device(t.device)
        self._grad_req = None
        self._grad_hook = None
        self._bw_epoch = -1

    # ---- basic properties -------------------------------------------------
    @property
    def data_(self):
        return self._t

    @property
    def _fresh_grad(self):
        """True when a backward pass has written this array's gradient since the flag was last cleared (the Trainer's
        stale-gradient check, ndarray.py:_fresh_grad of the reference)."""
        return bool(getattr(self._t, "_mx_fresh_grad", False))

    @_fresh_grad.setter
    def _fresh_grad(self, state):
        self._t._mx_fresh_grad = bool(state)

    def to_dlpack_for_read(self):
        from torch.utils import dlpack
        return dlpack.to_dlpack(self._t.detach())

    to_dlpack_for_write = to_dlpack_for_read

    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def size(self):
        return int(self._t.numel())

    @property
    def ndim(self):
        return self._t.dim()

    @property
    def dtype(self):
        return np_dtype(self._t.dtype)

    @property
    def context(self):
        return self._ctx

    ctx = context

    @property
    def stype(self):
        return "default"

    @property
    def T(self):
        return NDArray(self._t.t() if self._t.dim() == 2 else self._t.permute(*reversed(range(self._t.dim()))), self._ctx)

    @property
    def grad(self):
        g = self._t.grad
        if g is None:
            return None
        st = getattr(self, "_grad_stype", None)
        if st in ("row_sparse", "csr"):        # attach_grad(stype=...): the gradient is exposed with that storage type
            from .sparse import cast_storage
            return cast_storage(NDArray(g, self._ctx), st)
        return NDArray(g, self._ctx)

    def __repr__(self):
        return "\n%s\n<NDArray %s @%s>" % (str(self.asnumpy()), "x".join(str(s) for s in self.shape), self._ctx)

    def __len__(self):
        return self.shape[0]

    def __bool__(self):
        if self.size == 0:
            return False
        if self.size == 1:
            return bool(self._t.item())
        raise ValueError("The truth value of an NDArray with multiple elements is ambiguous.")

    __nonzero__ = __bool__

    # ---- conversion ------------------------------------------------------
    def asnumpy(self):
        _raise_deferred(self._t)
        t = self._t.detach()
        if t.dtype is torch.bfloat16:
            t = t.float()
        c = t.cpu()
        a = c.numpy()
        # always a private copy, like the reference (ndarray.py asnumpy -> MXNDArraySyncCopyToCPU): a CPU tensor's
        # numpy() aliases its storage and would change under the caller at the next in-place update
        return a.copy() if c.data_ptr() == self._t.data_ptr() else a

    def asscalar(self):
        if self.size != 1:
            raise ValueError("The current array is not a scalar")
        return self.asnumpy().reshape(-1)[0]

    def item(self):
        return self._t.item()

    def astype(self, dtype, copy=True):
        td = torch_dtype(dtype)
        if not copy and td == self._t.dtype:
            return self
        return _op(lambda t: t.to(td, copy=True), self)

    def copy(self):
        return _op(lambda t: t.clone(), self)

    def copyto(self, other):
        """Copy into another NDArray (in place) or onto a Context (new array)."""
        if isinstance(other, NDArray):
            if other is self:
                return other
            if getattr(other, "stype", "default") != "default":
                other[:] = self                     # dense -> sparse destination: storage conversion (cast_storage)
                return other
            with torch.no_grad():
                other._t.copy_(self._t, non_blocking=True)
            if getattr(self._t, "_mx_exc", None) is not None:
                other._t._mx_exc = self._t._mx_exc
            return other
        ctx = _ctx_of(other)
        with torch.no_grad():
            t = self._t.detach().to(ctx.torch_device, copy=True)
        if getattr(self._t, "_mx_exc", None) is not None:
            t._mx_exc = self._t._mx_exc
        return NDArray(_pin_if(t, ctx), ctx)

    def as_in_context(self, ctx):
        ctx = _ctx_of(ctx)
        if ctx == self._ctx:
            return self
        if ctx.torch_device == self._t.device:
            return NDArray(self._t, ctx)
        return self.copyto(ctx)

    def as_torch(self):
        return self._t

    def tostype(self, stype):
        if stype == "default":
            return self
        from . import sparse
        return sparse.cast_storage(self, stype)

    def wait_to_read(self):
        _raise_deferred(self._t)
        if self._t.is_cuda:
            torch.cuda.current_stream(self._t.device).synchronize()

    wait_to_write = wait_to_read

    def detach(self):
        return NDArray(self._t.detach(), self._ctx)

    # ---- autograd ---------------------------------------------------------
    def attach_grad(self, grad_req="write", stype=None):
        if grad_req == "null":
            self._t.requires_grad_(False)
            self._grad_req = "null"
            return
        if not self._t.is_leaf:
            self._t = self._t.detach()
        self._t.requires_grad_(True)
        if self._t.grad is None:
            self._t.grad = torch.zeros_like(self._t)
        self._grad_req = grad_req
        if stype is not None:
            self._grad_stype = stype
        _ag._register_leaf(self)

    def backward(self, out_grad=None, retain_graph=False, train_mode=True):
        _ag.backward([self], [out_grad] if out_grad is not None else None, retain_graph, train_mode)

    def zero_grad(self):
        if self._t.grad is not None:
            self._t.grad.zero_()

    # ---- indexing ----------------------------------------------------------
    def __getitem__(self, key):
        key = _unwrap_key(key)
        return _op(lambda t: t[key], self)

    def __iter__(self):
        for i in range(self.shape[0] if self.ndim else 0):
            yield self[i]

    def __setitem__(self, key, value):
        key = _unwrap_key(key)
        with torch.no_grad():
            v = _unwrap(value)
            if isinstance(v, (list, tuple)):
                v = np.asarray(v)
            if isinstance(v, np.generic):
                v = v.item()
            if isinstance(v, np.ndarray):
                v = torch.from_numpy(np.ascontiguousarray(v))
            if isinstance(v, torch.Tensor):
                v = v.to(device=self._t.device, dtype=self._t.dtype)
            if isinstance(key, slice) and key == slice(None):
                if isinstance(v, torch.Tensor):
                    if v.numel() == self._t.numel() and v.shape != self._t.shape:
                        v = v.reshape(self._t.shape)          # e.g. shape () <- (1,)
                    self._t.copy_(v.expand_as(self._t) if v.shape != self._t.shape else v)
                else:
                    self._t.fill_(v)
            else:
                self._t[key] = v

    # ---- shape ops -----------------------------------------------------------
    def reshape(self, *shape, **kwargs):
        if len(shape) == 1 and isinstance(shape[0], (list, tuple)):
            shape = tuple(shape[0])
        if not shape:
            shape = tuple(kwargs["shape"])
        if kwargs.get("reverse", False):
            from . import op as _o
            return _o.reshape(self, shape=shape, reverse=True)
        shape = _mx_reshape(self.shape, shape)
        return _op(lambda t: t.reshape(shape), self)

    def reshape_like(self, other=None, rhs=None, **kwargs):
        other = rhs if other is None else other
        if kwargs:                               # lhs_begin / lhs_end / rhs_begin / rhs_end ranges: the operator handles them
            from . import op as _op_mod
            return _op_mod.reshape_like(self, other, **kwargs)
        return self.reshape(other.shape)

    def flatten(self):
        return self.reshape((self.shape[0], -1))

    def expand_dims(self, axis):
        return _op(lambda t: t.unsqueeze(axis), self)

    def squeeze(self, axis=None):
        return _op(lambda t: t.squeeze() if axis is None else t.squeeze(axis), self)

    def transpose(self, *axes, **kwargs):
        if "axes" in kwargs:
            axes = (kwargs["axes"],) if kwargs["axes"] is not None else ()
        if len(axes) == 1 and isinstance(axes[0], (list, tuple)):
            axes = tuple(axes[0])
        if not axes:
            axes = tuple(reversed(range(self.ndim)))
        return _op(lambda t: t.permute(*axes), self)

    def swapaxes(self, a, b):
        return _op(lambda t: t.transpose(a, b), self)

    def broadcast_to(self, shape):
        return _op(lambda t: t.expand(*shape), self)

    def broadcast_like(self, other):
        return self.broadcast_to(other.shape)

    def tile(self, reps):
        return _op(lambda t: t.repeat(*reps), self)

    def slice_axis(self, axis, begin, end=None):
        def f(t):
            n = t.shape[axis]
            b = begin + n if begin < 0 else begin
            e = n if end is None else (end + n if end < 0 else end)
            return t.narrow(axis, b, e - b)
        return _op(f, self)

    def split(self, num_outputs, axis=1, squeeze_axis=False):
        from . import op
        return op.split(self, num_outputs=num_outputs, axis=axis, squeeze_axis=squeeze_axis)

    # ---- reductions / math (method forms) -----------------------------------
    def _reduce(self, fn, axis, keepdims, **kw):
        def f(t):
            if axis is None or axis == ():
                r = fn(t, **kw)
                return r.reshape((1,) * t.dim()) if keepdims else r.reshape(1) if r.dim() == 0 else r
            return fn(t, dim=axis, keepdim=keepdims, **kw)
        return _op(f, self)

    def sum(self, axis=None, keepdims=False):
        return self._reduce(torch.sum, axis, keepdims)

    def mean(self, axis=None, keepdims=False):
        return self._reduce(torch.mean, axis, keepdims)

    def prod(self, axis=None, keepdims=False):
        return self._reduce(torch.prod, axis, keepdims)

    def max(self, axis=None, keepdims=False):
        return self._reduce(torch.amax, axis if axis is not None else tuple(range(self.ndim)), keepdims) \
            if axis is not None else _op(lambda t: t.max().reshape(1), self)

    def min(self, axis=None, keepdims=False):
        return self._reduce(torch.amin, axis, keepdims) \
            if axis is not None else _op(lambda t: t.min().reshape(1), self)

    def norm(self, ord=2, axis=None, keepdims=False):
        return _op(lambda t: torch.linalg.vector_norm(t.float(), ord=ord, dim=axis, keepdim=keepdims).to(t.dtype).reshape(
            (1,) if axis is None and not keepdims else -1) if axis is None else
            torch.linalg.vector_norm(t.float(), ord=ord, dim=axis, keepdim=keepdims).to(t.dtype), self)

    def argmax(self, axis=None, keepdims=False):
        return _op(lambda t: torch.argmax(t, dim=axis, keepdim=keepdims).to(torch.float32), self)

    def argmin(self, axis=None, keepdims=False):
        return _op(lambda t: torch.argmin(t, dim=axis, keepdim=keepdims).to(torch.float32), self)

    def abs(self):
        return _op(torch.abs, self)

    def sign(self):
        return _op(torch.sign, self)

    def sqrt(self):
        return _op(torch.sqrt, self)

    def square(self):
        return _op(torch.square, self)

    def exp(self):
        return _op(torch.exp, self)

    def log(self):
        return _op(torch.log, self)

    def relu(self):
        return _op(torch.relu, self)

    def sigmoid(self):
        return _op(torch.sigmoid, self)

    def tanh(self):
        return _op(torch.tanh, self)

    def softmax(self, axis=-1):
        return _op(lambda t: torch.softmax(t, dim=axis), self)

    def log_softmax(self, axis=-1):
        return _op(lambda t: torch.log_softmax(t, dim=axis), self)

    def clip(self, a_min, a_max):
        return _op(lambda t: torch.clamp(t, a_min, a_max), self)

    def dot(self, other):
        from . import op
        return op.dot(self, other)

    def one_hot(self, depth, on_value=1.0, off_value=0.0, dtype="float32"):
        from . import op
        return op.one_hot(self, depth, on_value, off_value, dtype)

    def topk(self, **kw):
        from . import op
        return op.topk(self, **kw)

    def sort(self, axis=-1, is_ascend=True):
        return _op(lambda t: torch.sort(t, dim=axis, descending=not is_ascend)[0], self)

    def argsort(self, axis=-1, is_ascend=True):
        return _op(lambda t: torch.argsort(t, dim=axis, descending=not is_ascend).to(torch.float32), self)

    def flip(self, axis):
        return _op(lambda t: torch.flip(t, [axis]), self)

    def zeros_like(self):
        return zeros_like(self)

    def ones_like(self):
        return ones_like(self)

    # ---- arithmetic ---------------------------------------------------------
    def __add__(self, o):
        return _binop(torch.add, self, o)

    __radd__ = __add__

    def __sub__(self, o):
        return _binop(torch.sub, self, o)

    def __rsub__(self, o):
        return _binop(lambda a, b: torch.sub(b, a), self, o)

    def __mul__(self, o):
        return _binop(torch.mul, self, o)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return _binop(_mx_div, self, o)

    def __rtruediv__(self, o):
        return _binop(lambda a, b: _mx_div(b, a), self, o)

    __div__ = __truediv__
    __rdiv__ = __rtruediv__

    def __mod__(self, o):
        return _binop(torch.remainder, self, o)

    def __pow__(self, o):
        return _binop(torch.pow, self, o)

    def __rpow__(self, o):
        return _binop(lambda a, b: torch.pow(b, a), self, o)

    def __neg__(self):
        return _op(torch.neg, self)

    def __abs__(self):
        return _op(torch.abs, self)

    def _inplace(self, fn, o):
        with torch.no_grad():
            v = _unwrap(o)
            if isinstance(v, torch.Tensor) and v.device != self._t.device:
                v = v.to(self._t.device)
            fn(self._t, v)
        return self

    def __iadd__(self, o):
        return self._inplace(lambda t, v: t.add_(v), o)

    def __isub__(self, o):
        return self._inplace(lambda t, v: t.sub_(v), o)

    def __imul__(self, o):
        return self._inplace(lambda t, v: t.mul_(v), o)

    def __itruediv__(self, o):
        return self._inplace(lambda t, v: t.div_(v) if t.is_floating_point() else t.copy_(_mx_div(t, v)), o)

    __idiv__ = __itruediv__

    def _cmp(self, fn, o):
        return _binop(lambda a, b: fn(a, b).to(a.dtype if isinstance(a, torch.Tensor) else torch.float32), self, o)

    def __eq__(self, o):
        if o is None:
            return False
        return self._cmp(torch.eq, o)

    def __ne__(self, o):
        if o is None:
            return True
        return self._cmp(torch.ne, o)

    def __gt__(self, o):
        return self._cmp(torch.gt, o)

    def __ge__(self, o):
        return self._cmp(torch.ge, o)

    def __lt__(self, o):
        return self._cmp(torch.lt, o)

    def __le__(self, o):
        return self._cmp(torch.le, o)

    __hash__ = object.__hash__

    def __array__(self, dtype=None, copy=None):
        a = self.asnumpy()
        return a.astype(dtype) if dtype is not None else a

    def __getstate__(self):
        return {"data": self.asnumpy(), "bf16": self._t.dtype is torch.bfloat16}

    def __setstate__(self, st):
        t = torch.from_numpy(np.ascontiguousarray(st["data"]))
        if st.get("bf16"):
            t = t.to(torch.bfloat16)
        self._t = t
        self._ctx = Context("cpu", 0)
        self._grad_req = None
        self._grad_hook = None
        self._bw_epoch = -1


def _unwrap_key(key):
    if isinstance(key, NDArray):
        return key._t.long()
    if isinstance(key, tuple):
        return tuple(_unwrap_key(k) for k in key)
    return key


def _mx_div(a, b):
    """Division keeps the array's dtype: integer arrays divide like C (truncation), as the reference kernels do."""
    if isinstance(a, torch.Tensor) and not a.is_floating_point() and not (isinstance(b, torch.Tensor) and b.is_floating_point()) \
            and not isinstance(b, float):
        return torch.div(a, b, rounding_mode="trunc")
    if isinstance(a, torch.Tensor) and not a.is_floating_point() and isinstance(b, float):
        return torch.div(a, b, rounding_mode="trunc").to(a.dtype)
    return torch.div(a, b)


def _mx_reshape(src, shape, reverse=False):
    """MXNet reshape special codes: 0 copy dim, -1 infer, -2 copy rest, -3 merge two, -4 split.  ``reverse``: the codes
    are matched against the source dimensions from the right (matrix_op-inl.h ReshapeParam.reverse)."""
    if reverse:
        # the reference reverses the source dimensions AND the whole code vector, infers, and reverses the result: in this
        # mode a -4 is therefore written AFTER its two arguments ((16, -1, -4) splits the last dimension)
        out = _mx_reshape(tuple(reversed(src)), tuple(reversed(list(shape))), reverse=False)
        return tuple(reversed(out))
    if all(isinstance(s, (int, np.integer)) and s > 0 or s == -1 for s in shape):
        return tuple(int(s) for s in shape)
    out, i, j = [], 0, 0
    shape = list(shape)
    while j < len(shape):
        s = shape[j]
        if s == 0:
            out.append(src[i]); i += 1
        elif s == -1:
            out.append(-1); i += 1
        elif s == -2:
            out.extend(src[i:]); i = len(src)
        elif s == -3:
            out.append(src[i] * src[i + 1]); i += 2
        elif s == -4:
            a, b = shape[j + 1], shape[j + 2]
            if a == -1:
                a = src[i] // b
            if b == -1:
                b = src[i] // a
            out.extend([a, b]); i += 1; j += 2
        else:
            out.append(int(s)); i += 1
        j += 1
    return tuple(out)


def _op(fn, *arrs, **kw):
    """Run ``fn`` on unwrapped tensors honouring the autograd recording flag."""
    ctx = _result_ctx(*arrs)
    ts = [_unwrap(a) for a in arrs]
    try:
        if _ag.is_recording():
            r = fn(*ts, **kw)
        else:
            with torch.no_grad():
                r = fn(*ts, **kw)
    except MXNetError:
        raise
    except (RuntimeError, ValueError) as e:      # (IndexError stays: python's iteration protocol depends on it)
        # every operator failure is an MXNetError for the caller, whatever the backend raised (c_api_error.h)
        raise MXNetError("%s: %s" % (type(e).__name__, e)) from e
    for t in ts:
        e = getattr(t, "_mx_exc", None) if isinstance(t, torch.Tensor) else None
        if e is not None:                      # inputs carried a deferred failure: so do the outputs
            for o in (r if isinstance(r, (tuple, list)) else [r]):
                if isinstance(o, torch.Tensor):
                    o._mx_exc = e
            break
    return _wrap(r, ctx)


def _poison(t, exc):
    """Mark a tensor as the result of an operation that failed asynchronously: like the reference's engine, the error is
    raised when the value is needed on the host (asnumpy / wait_to_read / waitall), and it propagates to every array
    computed from this one (test_exc_handling.py semantics, threaded_engine.cc exception_ptr propagation)."""
    t._mx_exc = exc
    _PENDING.append(exc)
    return t


_PENDING = []


def _raise_deferred(t):
    e = getattr(t, "_mx_exc", None)
    if e is not None and not getattr(e, "_mx_delivered", False):
        if e in _PENDING:
            _PENDING.remove(e)
        # an operator's failure is delivered ONCE, at the first wait on any array that depends on it (threaded_engine.cc
        # rethrows and clears the exception of the variable); later reads of dependants return without raising
        e._mx_delivered = True
        raise e


def _wrap(r, ctx):
    if isinstance(r, torch.Tensor):
        return NDArray(r, ctx)
    if isinstance(r, (tuple, list)):
        return [_wrap(x, ctx) for x in r]
    return r


def _binop(fn, a, b):
    if isinstance(b, NDArray):
        tb = b._t
        if tb.device != a._t.device:
            raise MXNetError("operands on different devices: %s vs %s" % (a._ctx, b._ctx))
        return _op(fn, a, b)
    if isinstance(b, np.ndarray):
        b = torch.from_numpy(b).to(a._t.device, a._t.dtype)
    elif isinstance(b, np.generic):
        b = b.item()
    return _op(lambda t: fn(t, b), a)


# ---------------------------------------------------------------------------
# creation
# ---------------------------------------------------------------------------

def from_torch(t, ctx=None):
    return NDArray(t, ctx)


def array(source, ctx=None, dtype=None):
    # sparse sources keep their storage type (ndarray/utils.py:array dispatches scipy.sparse / sparse NDArrays to sparse.array)
    if getattr(source, "stype", "default") != "default" or type(source).__module__.startswith("scipy.sparse"):
        from . import sparse as _sp
        return _sp.array(source, ctx=ctx, dtype=dtype)
    ctx = _ctx_of(ctx)
    if isinstance(source, NDArray):
        t = source._t.detach()
        td = torch_dtype(dtype) if dtype is not None else t.dtype
        return NDArray(_pin_if(t.to(device=ctx.torch_device, dtype=td, copy=True), ctx), ctx)
    if isinstance(source, torch.Tensor):
        td = torch_dtype(dtype) if dtype is not None else source.dtype
        return NDArray(_pin_if(source.detach().to(device=ctx.torch_device, dtype=td, copy=True), ctx), ctx)
    a = np.asarray(source)
    if dtype is None:
        # anything that is not already an NDArray becomes float32 unless a dtype is given (ndarray/utils.py:array of the
        # reference: ``dtype = mx_real_t if dtype is None``) -- integer numpy arrays included
        dtype = np.float32
    td = torch_dtype(dtype)
    t = torch.from_numpy(np.array(a, copy=True, order="C")).to(td)   # always own the memory (never alias numpy)
    if t.dim() == 0 and not isinstance(source, np.ndarray):
        t = t.reshape(1)        # python scalars become 1-element arrays; a 0-d numpy array keeps shape ()
    return NDArray(_pin_if(t.to(ctx.torch_device), ctx), ctx)


def _shape(shape):
    return (int(shape),) if isinstance(shape, (int, np.integer)) else tuple(int(s) for s in shape)


def _new(fn, shape, ctx, dtype, **kw):
    ctx = _ctx_of(ctx)
    pin = ctx.device_typeid == 3 and torch.cuda.is_available()
    t = fn(_shape(shape), dtype=torch_dtype(dtype), device=ctx.torch_device, pin_memory=pin, **kw)
    return NDArray(t, ctx)


def _into(out, res):
    if out is None:
        return res
    out[:] = res
    return out


def zeros(shape, ctx=None, dtype=None, stype=None, out=None, **kw):
    if stype in ("row_sparse", "csr"):
        from . import sparse
        return sparse.zeros(stype, shape, ctx=ctx, dtype=dtype)
    if out is not None:
        out[:] = 0
        return out
    return _new(torch.zeros, shape, ctx, dtype)


def ones(shape, ctx=None, dtype=None, out=None, **kw):
    if out is not None:
        out[:] = 1
        return out
    return _new(torch.ones, shape, ctx, dtype)


def from_dlpack(capsule):
    """NDArray sharing memory with a DLPack capsule (or any object with ``__dlpack__``)."""
    from torch.utils import dlpack
    return NDArray(dlpack.from_dlpack(capsule))


def to_dlpack_for_read(data):
    return data.to_dlpack_for_read()


def to_dlpack_for_write(data):
    return data.to_dlpack_for_write()


def empty(shape, ctx=None, dtype=None, stype=None):
    if stype not in (None, "default"):
        from . import sparse as _sp
        return _sp.zeros(stype, shape, ctx=ctx, dtype=dtype)
    return _new(torch.empty, shape, ctx, dtype)


def full(shape, val, ctx=None, dtype=None, out=None):
    r = _new(torch.empty, shape, ctx, dtype)
    r._t.fill_(val)
    if out is not None:
        out[:] = r
        return out
    return r


def arange(start, stop=None, step=1.0, repeat=1, ctx=None, dtype=None):
    ctx = _ctx_of(ctx)
    if stop is None:
        start, stop = 0, start
    t = torch.arange(start, stop, step, dtype=torch_dtype(dtype), device=ctx.torch_device)
    if repeat > 1:
        t = t.repeat_interleave(repeat)
    return NDArray(t, ctx)


def zeros_like(a):
    return NDArray(torch.zeros_like(a._t), a._ctx)


def ones_like(a):
    return NDArray(torch.ones_like(a._t), a._ctx)


def concatenate(arrays, axis=0, always_copy=True):
    return _op(lambda *ts: torch.cat(ts, dim=axis), *arrays)


def moveaxis(a, source, destination):
    """Move one axis to a new position (ndarray.py:moveaxis of the reference: list pop / insert, so a destination past the
    end means "last")."""
    axes = list(range(a.ndim))
    try:
        axes.pop(source)
    except IndexError:
        raise ValueError("Source should verify 0 <= source < tensor.ndim. Got %d." % source)
    try:
        axes.insert(destination, source)
    except IndexError:
        raise ValueError("Destination should verify 0 <= destination < tensor.ndim. Got %d." % destination)
    return _op(lambda t: t.permute(*axes), a)


def waitall():
    """Block until all queued device work is done (MXNDArrayWaitAll, c_api.cc:287)."""
    from .. import engine
    engine.get().wait_all()
    if torch.cuda.is_available():
        for d in range(torch.cuda.device_count()):
            torch.cuda.synchronize(d)
    if _PENDING:
        e = _PENDING.pop(0)
        del _PENDING[:]
        raise e
