"""Sparse storage types: row_sparse and csr.

Reference: python/mxnet/ndarray/sparse.py, include/mxnet/ndarray.h (aux data
per storage type), src/operator/tensor/cast_storage-inl.cuh, dot-inl.cuh.
A RowSparseNDArray is (data[nnz_rows, ...], indices[nnz_rows]) with sorted
unique int64 row ids; a CSRNDArray is (data, indices, indptr).
"""
import numpy as np
import warnings

import torch

from ..base import MXNetError

warnings.filterwarnings("ignore", message="Sparse CSR tensor support is in beta state")
warnings.filterwarnings("ignore", message="Sparse invariant checks are implicitly disabled")

from ..base import torch_dtype, np_dtype
from ..context import current_context
from .ndarray import NDArray, _ctx_of, array as _dense_array

__all__ = ["BaseSparseNDArray", "RowSparseNDArray", "CSRNDArray", "row_sparse_array", "csr_matrix",
           "zeros", "cast_storage", "dot", "retain", "add_n_rsp"]


class BaseSparseNDArray(NDArray):
    """Sparse arrays are NDArrays (python/mxnet/ndarray/sparse.py:BaseSparseNDArray): any operator or method without a
    sparse implementation sees the dense view ``_t`` (the reference's storage fallback) and returns a dense result."""
    stype = None
    _grad_req = None
    _grad_hook = None
    _bw_epoch = -1

    @property
    def context(self):
        return self._ctx

    ctx = context

    @property
    def dtype(self):
        return np_dtype(self._data.dtype)

    @property
    def size(self):
        return int(np.prod(self._shape))

    @property
    def ndim(self):
        return len(self._shape)

    @property
    def shape(self):
        return tuple(self._shape)

    def asnumpy(self):
        return self.todense().asnumpy()

    def asscalar(self):
        return self.todense().asscalar()

    @property
    def _t(self):
        """Dense torch view for operators without a sparse kernel (the reference's storage fallback).  READ ONLY: writes
        go through ``x[:] = value`` / ``copyto``.  After ``attach_grad()`` the dense leaf that owns the gradient is returned, so
        operators recorded by autograd are connected to it."""
        leaf = self.__dict__.get("_dense_leaf")
        return leaf._t if leaf is not None else self.todense()._t

    @_t.setter
    def _t(self, value):
        self[:] = NDArray(value, self._ctx)

    @property
    def T(self):
        return self.todense().T

    def wait_to_write(self):
        self.wait_to_read()

    def zeros_like(self):
        return zeros(self.stype, self.shape, ctx=self._ctx, dtype=self.dtype)

    def reshape(self, *a, **k):
        from ..base import NotSupportedForSparseNDArray
        raise NotSupportedForSparseNDArray(self.reshape, None, *a)

    # ---- autograd: gradients of sparse leaves are kept dense-backed and exposed with the requested storage type
    def attach_grad(self, grad_req="write", stype=None):
        self._grad_stype = stype or self.stype
        self._dense_leaf = self.todense()
        self._dense_leaf.attach_grad(grad_req)

    @property
    def grad(self):
        leaf = getattr(self, "_dense_leaf", None)
        if leaf is None or leaf.grad is None:
            return None
        st = getattr(self, "_grad_stype", "default")
        return cast_storage(leaf.grad, st) if st != "default" else leaf.grad

    def detach(self):
        return self.copy()

    def __neg__(self):
        return self * -1.0

    def __radd__(self, o):
        return self.__add__(o)

    def __rmul__(self, o):
        return self.__mul__(o)

    def __iadd__(self, o):
        r = self + o
        self[:] = r
        return self

    def __isub__(self, o):
        r = self - o
        self[:] = r
        return self

    def __imul__(self, o):
        r = self * o
        self[:] = r
        return self

    def __itruediv__(self, o):
        r = self / o
        self[:] = r
        return self

    def sum(self, *a, **k):
        return self.todense().sum(*a, **k)

    def mean(self, *a, **k):
        return self.todense().mean(*a, **k)

    def square(self):
        return self * self

    def _sync_leaf(self):
        """After a write into a sparse array that carries a gradient: refresh the dense leaf operators read."""
        leaf = self.__dict__.get("_dense_leaf")
        if leaf is not None:
            with torch.no_grad():
                leaf._t.copy_(self.todense()._t)

    def check_format(self, full_check=True):
        """Raise MXNetError if the index arrays are malformed (ndarray/sparse.py check_format -> common/utils.h
        CheckFormatWrapper): csr -- indptr starts at 0, is non-decreasing, ends at nnz, column ids within [0, cols) and
        ascending inside a row; row_sparse -- row ids within [0, rows) and strictly ascending."""
        from ..base import MXNetError
        idx = self._indices
        if self.stype == "row_sparse":
            if idx.numel() and (bool((idx < 0).any()) or bool((idx >= self._shape[0]).any()) or bool((idx[1:] <= idx[:-1]).any())):
                raise MXNetError("row_sparse indices must be within [0, %d) and strictly ascending" % self._shape[0])
            return
        ptr = self._indptr
        if ptr.numel() != self._shape[0] + 1 or int(ptr[0]) != 0 or int(ptr[-1]) != idx.numel() or bool((ptr[1:] < ptr[:-1]).any()):
            raise MXNetError("csr indptr must start at 0, be non-decreasing and end at the number of stored values")
        if idx.numel() and (bool((idx < 0).any()) or bool((idx >= self._shape[1]).any())):
            raise MXNetError("csr indices must be within [0, %d)" % self._shape[1])
        if full_check and idx.numel() > 1:
            rows = torch.repeat_interleave(torch.arange(self._shape[0], device=ptr.device), ptr[1:] - ptr[:-1])
            same_row = rows[1:] == rows[:-1]
            if bool(((idx[1:] <= idx[:-1]) & same_row).any()):
                raise MXNetError("csr indices must be ascending inside every row")

    @property
    def _num_aux(self):
        return 1 if self.stype == "row_sparse" else 2

    def __getstate__(self):
        st = {"stype": self.stype, "shape": self._shape, "data": self._data.cpu().numpy(), "indices": self._indices.cpu().numpy()}
        if self.stype == "csr":
            st["indptr"] = self._indptr.cpu().numpy()
        return st

    def __setstate__(self, st):
        self._ctx = current_context()
        self._shape = tuple(st["shape"])
        self._data = torch.from_numpy(st["data"])
        self._indices = torch.from_numpy(st["indices"]).to(torch.int64)
        if st["stype"] == "csr":
            self._indptr = torch.from_numpy(st["indptr"]).to(torch.int64)

    def wait_to_read(self):
        if self._data.is_cuda:
            torch.cuda.current_stream(self._data.device).synchronize()

    def tostype(self, stype):
        if stype == self.stype:
            return self
        return cast_storage(self.todense(), stype) if stype != "default" else self.todense()

    def __repr__(self):
        return "\n<%s %s @%s>" % (type(self).__name__, "x".join(map(str, self.shape)), self._ctx)


def _recorded_dense(a, b, fn, ctx):
    """Arithmetic on sparse arrays under ``autograd.record()``: when an operand carries a gradient (attach_grad /
    mark_variables) the operation runs on the dense leaves so that it is taped -- the reference's storage fallback for
    operators without a sparse backward.  Returns None when nothing needs recording."""
    from .. import autograd as _ag
    if not _ag.is_recording():
        return None
    has_leaf = lambda v: isinstance(v, BaseSparseNDArray) and v.__dict__.get("_dense_leaf") is not None   # noqa: E731
    needs = has_leaf(a) or has_leaf(b) or (isinstance(b, NDArray) and not isinstance(b, BaseSparseNDArray) and b._t.requires_grad)
    if not needs:
        return None
    tb = b._t if isinstance(b, NDArray) else b
    return NDArray(fn(a._t, tb), ctx)



class RowSparseNDArray(BaseSparseNDArray):
    stype = "row_sparse"

    def __init__(self, data, indices, shape, ctx=None):
        self._data = data
        self._indices = indices.to(torch.int64)
        self._shape = tuple(int(s) for s in shape)
        self._ctx = ctx if ctx is not None else current_context()

    @property
    def data(self):
        return NDArray(self._data, self._ctx)

    @property
    def indices(self):
        return NDArray(self._indices, self._ctx)

    def todense(self):
        out = torch.zeros(self._shape, dtype=self._data.dtype, device=self._data.device)
        if self._indices.numel():
            out.index_copy_(0, self._indices, self._data)
        return NDArray(out, self._ctx)

    def copy(self):
        return RowSparseNDArray(self._data.clone(), self._indices.clone(), self._shape, self._ctx)

    def copyto(self, other):
        if isinstance(other, RowSparseNDArray):
            other._data = self._data.to(other._data.device, copy=True)
            other._indices = self._indices.to(other._data.device, copy=True)
            other._shape = self._shape
            other._sync_leaf()
            return other
        if isinstance(other, NDArray):
            other[:] = self.todense()
            return other
        ctx = _ctx_of(other)
        return RowSparseNDArray(self._data.to(ctx.torch_device, copy=True), self._indices.to(ctx.torch_device, copy=True),
                                self._shape, ctx)

    def as_in_context(self, ctx):
        return self if ctx == self._ctx else self.copyto(ctx)

    def astype(self, dtype, copy=True):
        if not copy and torch_dtype(dtype) == self._data.dtype:
            return self
        return RowSparseNDArray(self._data.to(torch_dtype(dtype), copy=True), self._indices.clone(), self._shape, self._ctx)

    def retain(self, row_ids):
        return retain(self, row_ids)

    def __setitem__(self, key, value):
        assert isinstance(key, slice) and key == slice(None)
        if isinstance(value, RowSparseNDArray):
            value.copyto(self)
        else:
            src = cast_storage(value if isinstance(value, NDArray) else _dense_array(value, ctx=self._ctx), "row_sparse")
            src.copyto(self)

    def _binary(self, other, fn):
        rec = _recorded_dense(self, other, fn, self._ctx)
        if rec is not None:
            return rec
        if isinstance(other, RowSparseNDArray):
            return cast_storage(NDArray(fn(self.todense()._t, other.todense()._t), self._ctx), "row_sparse")
        if isinstance(other, NDArray):
            return NDArray(fn(self.todense()._t, other._t), self._ctx)
        return RowSparseNDArray(fn(self._data, other), self._indices, self._shape, self._ctx)

    def __add__(self, o):
        return self._binary(o, torch.add)

    def __sub__(self, o):
        return self._binary(o, torch.sub)

    def __mul__(self, o):
        return self._binary(o, torch.mul)

    def __truediv__(self, o):
        return self._binary(o, torch.div)


class CSRNDArray(BaseSparseNDArray):
    stype = "csr"

    def __init__(self, data, indices, indptr, shape, ctx=None):
        self._data = data
        self._indices = indices.to(torch.int64)
        self._indptr = indptr.to(torch.int64)
        self._shape = tuple(int(s) for s in shape)
        self._ctx = ctx if ctx is not None else current_context()

    @property
    def data(self):
        return NDArray(self._data, self._ctx)

    @property
    def indices(self):
        return NDArray(self._indices, self._ctx)

    @property
    def indptr(self):
        return NDArray(self._indptr, self._ctx)

    def _torch_csr(self):
        return torch.sparse_csr_tensor(self._indptr, self._indices, self._data, size=self._shape)

    def todense(self):
        return NDArray(self._torch_csr().to_dense(), self._ctx)

    def copy(self):
        return CSRNDArray(self._data.clone(), self._indices.clone(), self._indptr.clone(), self._shape, self._ctx)

    def copyto(self, other):
        if isinstance(other, CSRNDArray):
            other._data, other._indices, other._indptr, other._shape = self._data.clone(), self._indices.clone(), self._indptr.clone(), self._shape
            other._sync_leaf()
            return other
        if isinstance(other, NDArray):
            other[:] = self.todense()
            return other
        ctx = _ctx_of(other)
        d = ctx.torch_device
        return CSRNDArray(self._data.to(d, copy=True), self._indices.to(d, copy=True), self._indptr.to(d, copy=True), self._shape, ctx)

    def as_in_context(self, ctx):
        return self if ctx == self._ctx else self.copyto(ctx)

    def astype(self, dtype, copy=True):
        if not copy and torch_dtype(dtype) == self._data.dtype:
            return self
        return CSRNDArray(self._data.to(torch_dtype(dtype), copy=True), self._indices.clone(), self._indptr.clone(), self._shape,
                          self._ctx)

    def __setitem__(self, key, value):
        assert isinstance(key, slice) and key == slice(None), "CSRNDArray only supports [:] assignment"
        if isinstance(value, CSRNDArray):
            value.copyto(self)
        else:
            src = cast_storage(value if isinstance(value, NDArray) else _dense_array(value, ctx=self._ctx), "csr")
            src.astype(self._data.dtype).copyto(self)

    def _binary(self, other, fn):
        rec = _recorded_dense(self, other, fn, self._ctx)
        if rec is not None:
            return rec
        if isinstance(other, BaseSparseNDArray):
            return cast_storage(NDArray(fn(self.todense()._t, other.todense()._t), self._ctx), "csr")
        if isinstance(other, NDArray):
            return NDArray(fn(self.todense()._t, other._t), self._ctx)
        return CSRNDArray(fn(self._data, other), self._indices, self._indptr, self._shape, self._ctx)

    def __add__(self, o):
        return self._binary(o, torch.add)

    def __sub__(self, o):
        return self._binary(o, torch.sub)

    def __mul__(self, o):
        return self._binary(o, torch.mul)

    def __truediv__(self, o):
        return self._binary(o, torch.div)

    def asscipy(self):
        import scipy.sparse as spsp
        return spsp.csr_matrix((self._data.cpu().numpy(), self._indices.cpu().numpy(), self._indptr.cpu().numpy()), shape=self._shape)

    def __getitem__(self, key):
        if isinstance(key, int):
            key = slice(key, key + 1)
        assert isinstance(key, slice) and key.step in (None, 1)
        b, e, _ = key.indices(self._shape[0])
        lo, hi = int(self._indptr[b]), int(self._indptr[e])
        return CSRNDArray(self._data[lo:hi], self._indices[lo:hi], self._indptr[b:e + 1] - lo, (e - b, self._shape[1]), self._ctx)


def _src_dtype(src, dtype):
    """Default dtype of the sparse constructors: the source's own dtype for numpy / NDArray / scipy inputs, float32 for python
    lists (ndarray/sparse.py:_prepare_default_dtype of the reference)."""
    if dtype is not None:
        return dtype
    if isinstance(src, tuple) and src:
        src = src[0]
    if isinstance(src, (np.ndarray, NDArray)) or type(src).__module__.startswith("scipy.sparse"):
        return src.dtype
    return np.float32


def array(source_array, ctx=None, dtype=None):
    """Sparse array from a scipy.sparse matrix or another sparse NDArray (ndarray/sparse.py:array of the reference)."""
    if isinstance(source_array, BaseSparseNDArray):
        out = source_array.copyto(_ctx_of(ctx))
        return out if dtype is None else out.astype(dtype)
    if type(source_array).__module__.startswith("scipy.sparse"):
        return csr_matrix(source_array, ctx=ctx, dtype=dtype)
    raise ValueError("Unexpected source_array type: %s" % type(source_array))


def _is_shape(v):
    return isinstance(v, tuple) and len(v) > 0 and all(isinstance(x, (int, np.integer)) for x in v)


def row_sparse_array(arg1, shape=None, ctx=None, dtype=None):
    ctx = _ctx_of(ctx)
    if _is_shape(arg1):                               # row_sparse_array((D0, D1, ..)): an empty array of that shape
        if shape is not None and tuple(int(x) for x in shape) != tuple(int(x) for x in arg1):
            raise ValueError("shape %s does not match the shape given as first argument %s" % (tuple(shape), arg1))
        return zeros("row_sparse", tuple(int(x) for x in arg1), ctx=ctx, dtype=dtype or np.float32)
    dtype = _src_dtype(arg1, dtype)
    if isinstance(arg1, tuple) and len(arg1) == 2:
        data, idx = arg1
        data = _dense_array(data, ctx=ctx, dtype=dtype)._t
        idx = _dense_array(idx, ctx=ctx, dtype="int64")._t
        if shape is None:
            if idx.numel() == 0:
                raise ValueError("invalid shape: it cannot be inferred from empty indices")
            shape = (int(idx.max()) + 1,) + tuple(data.shape[1:])
        return RowSparseNDArray(data, idx, shape, ctx)
    if isinstance(arg1, RowSparseNDArray):
        return arg1.copyto(ctx).astype(dtype, copy=False)
    dense = _dense_array(arg1, ctx=ctx, dtype=dtype)
    if shape is not None and tuple(shape) != dense.shape:
        raise ValueError("shape %s does not match the source array's %s" % (tuple(shape), dense.shape))
    return cast_storage(dense, "row_sparse")


def csr_matrix(arg1, shape=None, ctx=None, dtype=None):
    ctx = _ctx_of(ctx)
    if _is_shape(arg1) and len(arg1) == 2:            # csr_matrix((M, N)): empty matrix
        if shape is not None and tuple(int(x) for x in shape) != tuple(int(x) for x in arg1):
            raise ValueError("shape %s does not match the shape given as first argument %s" % (tuple(shape), arg1))
        return zeros("csr", tuple(int(x) for x in arg1), ctx=ctx, dtype=dtype or np.float32)
    dtype = _src_dtype(arg1, dtype)
    if isinstance(arg1, tuple) and len(arg1) == 3:
        data, indices, indptr = arg1
        data = _dense_array(data, ctx=ctx, dtype=dtype)._t
        indices = _dense_array(indices, ctx=ctx, dtype="int64")._t
        indptr = _dense_array(indptr, ctx=ctx, dtype="int64")._t
        if shape is None:
            if indices.numel() == 0:
                raise ValueError("invalid shape: it cannot be inferred from empty indices")
            shape = (indptr.numel() - 1, int(indices.max()) + 1)
        return CSRNDArray(data, indices, indptr, shape, ctx)
    if type(arg1).__module__.startswith("scipy.sparse"):
        m = arg1.tocsr().copy()
        m.sum_duplicates()                           # canonical form: no duplicate entries, sorted column ids
        m.sort_indices()
        return csr_matrix((m.data, m.indices, m.indptr), shape=shape or m.shape, ctx=ctx, dtype=dtype or m.dtype)
    if isinstance(arg1, CSRNDArray):
        return arg1.copyto(ctx) if dtype is None else arg1.copyto(ctx).astype(dtype)
    if isinstance(arg1, tuple) and len(arg1) == 2:
        # (data, (row, col)) coordinate form
        import scipy.sparse as spsp
        data, (row, col) = arg1
        as_np = lambda v: v.asnumpy() if hasattr(v, "asnumpy") else np.asarray(v)   # noqa: E731
        return csr_matrix(spsp.coo_matrix((as_np(data), (as_np(row), as_np(col))), shape=shape), ctx=ctx, dtype=dtype)
    dense = _dense_array(arg1, ctx=ctx, dtype=dtype)
    if shape is not None and tuple(shape) != dense.shape:
        raise ValueError("shape %s does not match the source array's %s" % (tuple(shape), dense.shape))
    return cast_storage(dense, "csr")


def zeros(stype, shape, ctx=None, dtype=None):
    ctx = _ctx_of(ctx)
    d = ctx.torch_device
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    td = torch_dtype(dtype)
    if stype == "row_sparse":
        return RowSparseNDArray(torch.zeros((0,) + shape[1:], dtype=td, device=d), torch.zeros(0, dtype=torch.int64, device=d), shape, ctx)
    if stype == "csr":
        return CSRNDArray(torch.zeros(0, dtype=td, device=d), torch.zeros(0, dtype=torch.int64, device=d),
                          torch.zeros(shape[0] + 1, dtype=torch.int64, device=d), shape, ctx)
    if stype != "default":
        raise ValueError("unknown storage type " + str(stype))
    from .ndarray import zeros as dz
    return dz(shape, ctx, dtype)


def cast_storage(arr, stype):
    """src/operator/tensor/cast_storage-inl.cuh."""
    if isinstance(arr, BaseSparseNDArray):
        return arr.tostype(stype)
    t = arr._t
    if stype == "default":
        return arr
    if stype == "row_sparse":
        nz = (t.reshape(t.shape[0], -1) != 0).any(dim=1)
        idx = torch.nonzero(nz).reshape(-1)
        return RowSparseNDArray(t[idx].clone(), idx, t.shape, arr._ctx)
    if stype == "csr":
        c = t.to_sparse_csr()
        return CSRNDArray(c.values(), c.col_indices(), c.crow_indices(), t.shape, arr._ctx)
    raise ValueError("unknown stype " + str(stype))


def retain(data=None, indices=None, out=None, **kwargs):
    """sparse_retain: keep only the requested rows (src/operator/tensor/sparse_retain.cu).  ``retain(rsp, row_ids)`` or the
    operator's keyword form ``retain(data=rsp, indices=row_ids)``."""
    if kwargs:
        raise MXNetError("sparse_retain: unknown argument(s) %s" % sorted(kwargs))
    rsp, row_ids = data, indices
    if not isinstance(rsp, RowSparseNDArray):
        raise MXNetError("sparse retain expects a row_sparse input, got %s" % type(rsp).__name__)
    ids = row_ids._t.long() if isinstance(row_ids, NDArray) else torch.as_tensor(row_ids, dtype=torch.int64, device=rsp._data.device)
    from ..kvstore.utils import unique_rows
    ids = unique_rows(ids, num_rows=rsp._shape[0])
    mask = torch.isin(rsp._indices, ids)
    return RowSparseNDArray(rsp._data[mask], rsp._indices[mask], rsp._shape, rsp._ctx)


def add_n_rsp(arrs):
    """Sum of row_sparse arrays -> row_sparse (ElementWiseRspAdditionKernel, ndarray_function-inl.cuh)."""
    shape, ctx = arrs[0]._shape, arrs[0]._ctx
    idx = torch.cat([a._indices for a in arrs])
    dat = torch.cat([a._data for a in arrs])
    if idx.numel() == 0:
        return arrs[0].copy()
    uniq, inv = torch.unique(idx, return_inverse=True)
    out = torch.zeros((uniq.numel(),) + tuple(dat.shape[1:]), dtype=dat.dtype, device=dat.device)
    out.index_add_(0, inv, dat)
    return RowSparseNDArray(out, uniq, shape, ctx)


def dot(lhs, rhs, transpose_a=False, transpose_b=False, forward_stype=None, out=None):
    """Sparse dot: csr x dense, csr^T x dense -> dense (dot-inl.cuh); ``forward_stype`` asks for the result in that storage."""
    r = _dot(lhs, rhs, transpose_a, transpose_b)
    if forward_stype not in (None, "default"):
        r = cast_storage(r, forward_stype)
    if out is not None:
        out[:] = r
        return out
    return r


def _dot(lhs, rhs, transpose_a=False, transpose_b=False):
    has_leaf = lambda v: isinstance(v, BaseSparseNDArray) and v.__dict__.get("_dense_leaf") is not None   # noqa: E731
    if isinstance(lhs, CSRNDArray) and not has_leaf(lhs):
        a = lhs._torch_csr()
        b = rhs._t                      # (dense view / gradient leaf of a sparse right-hand side)
        if transpose_b:
            b = b.t()
        if transpose_a:
            return NDArray(torch.matmul(a.to_dense().t(), b), lhs._ctx)
        return NDArray(torch.matmul(a, b), lhs._ctx)
    a, b = lhs._t, rhs._t
    if transpose_a:
        a = a.t()
    if transpose_b:
        b = b.t()
    return NDArray(torch.matmul(a, b), lhs._ctx)


# ---- result storage types of operators that ran on the dense view (FInferStorageType of the reference registrations) ----
_ZERO_PRESERVING = {"relu", "abs", "sign", "sqrt", "square", "sin", "tan", "arcsin", "arctan", "sinh", "tanh", "arcsinh", "arctanh",
                    "floor", "ceil", "round", "rint", "fix", "trunc", "negative", "expm1", "log1p", "degrees", "radians", "cbrt",
                    "_mul_scalar", "_div_scalar", "zeros_like", "identity", "_copy", "stop_gradient", "BlockGrad", "make_loss"}
_SPARSE_ONLY_IF_ALL = {"elemwise_add", "elemwise_sub", "_plus", "_minus", "_grad_add", "add_n", "ElementWiseSum", "_add", "_sub",
                       "_maximum", "_minimum", "maximum", "minimum", "_hypot", "hypot"}
_SPARSE_IF_ANY = {"elemwise_mul", "_mul"}


def storage_rule(op, args, kwargs, result):
    """Storage type of ``result`` (computed densely) given the operator and its inputs; the value is unchanged.  Under
    autograd the sparse result keeps the taped dense tensor as its ``_dense_leaf`` so that consumers stay connected."""
    if isinstance(result, (list, tuple)) or isinstance(result, BaseSparseNDArray) or not isinstance(result, NDArray) or result.ndim < 1:
        return result
    arrs = [a for a in args if isinstance(a, NDArray)]
    sp = [a for a in arrs if isinstance(a, BaseSparseNDArray)]
    st = None
    if op in _ZERO_PRESERVING:
        st = sp[0].stype
    elif op in _SPARSE_IF_ANY:
        st = "row_sparse" if any(a.stype == "row_sparse" for a in sp) and not all(a.stype == "csr" for a in sp) else sp[0].stype
        if len(sp) == 2 and sp[0].stype != sp[1].stype:
            st = None
    elif op in _SPARSE_ONLY_IF_ALL:
        st = sp[0].stype if len(sp) == len(arrs) and all(a.stype == sp[0].stype for a in sp) else None
    elif op and op.startswith("_scatter_"):
        st = arrs[0].stype if isinstance(arrs[0], BaseSparseNDArray) else None
    elif op in ("square_sum", "_square_sum"):
        ax = kwargs.get("axis")
        ax = ax[0] if isinstance(ax, (list, tuple)) and len(ax) == 1 else ax
        st = "row_sparse" if sp[0].stype == "row_sparse" and ax in (1, -1) and kwargs.get("keepdims", False) else None
    elif op == "_contrib_quadratic":
        st = sp[0].stype if float(kwargs.get("c", 0.0)) == 0.0 and sp[0].stype == "csr" else None
    if st in (None, "default") or result.shape != sp[0].shape and op not in ("square_sum", "_square_sum"):
        return result
    if op and op.startswith("_scatter_") and st == "row_sparse":
        # scatter operators touch exactly the rows the left operand stores (elemwise_scatter_op.cc)
        idx = arrs[0]._indices
        return RowSparseNDArray(result._t.detach()[idx], idx.clone(), result.shape, result.context)
    if op and op.startswith("_scatter_") and st == "csr":
        src = arrs[0]                                  # ... and exactly the stored elements of a csr left operand
        rows = torch.repeat_interleave(torch.arange(src._shape[0], device=src._indptr.device), src._indptr[1:] - src._indptr[:-1])
        return CSRNDArray(result._t.detach()[rows, src._indices], src._indices.clone(), src._indptr.clone(), result.shape, result.context)
    if result.ndim != 2 and st == "csr":
        return result
    out = cast_storage(NDArray(result._t.detach(), result.context), st)
    if result._t.requires_grad:
        out._dense_leaf = result
    return out


def __getattr__(name):
    """mx.nd.sparse.<op>: operators without a dedicated sparse kernel fall back to the dense implementation on the dense
    view of their inputs (the reference generates this namespace from the same registry: ndarray/sparse.py + gen_sparse)."""
    if name.startswith("__"):
        raise AttributeError(name)
    from . import op as _opmod
    for cand in ("_sparse_" + name, name, {"add": "elemwise_add", "subtract": "elemwise_sub", "multiply": "elemwise_mul",
                                            "divide": "elemwise_div"}.get(name, "")):
        f = _opmod._REG.get(cand)
        if f is not None:
            return f
    raise AttributeError("module 'mxb200.ndarray.sparse' has no attribute %r" % name)
