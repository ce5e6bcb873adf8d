"""Testing helpers with the reference's names and call conventions (python/mxnet/test_utils.py), so operator / module test
scripts written against it run here: numeric-gradient checks, symbolic forward / backward checks against numpy
expectations, cross-context / cross-dtype consistency (`check_consistency`, the idea behind this repo's own GPU numerics
tests), random array factories incl. sparse ones, statistical generator checks, small utilities.

``check_numeric_gradient`` and ``check_consistency`` accept either a Symbol (reference convention) or a plain python
function of NDArrays (used by this repo's tests)."""
import contextlib
import functools
import numbers
import os
import sys
import time

import numpy as np
import torch

from . import ndarray as nd
from .base import MXNetError
from .context import Context, cpu, gpu, num_gpus, current_context
from .ndarray import NDArray

_DEFAULT_CTX = [None]
_RNG = np.random
# names that scripts get through ``from mxnet.test_utils import *`` in the reference (its module-level imports)
import numpy.random as rnd  # noqa: E402,F401
from .ndarray import array, zeros, ones  # noqa: E402,F401
from .symbol import Symbol  # noqa: E402,F401
import mxb200 as mx  # noqa: E402,F401  (the package itself; safe: it is already in sys.modules when this module loads)


# ----------------------------------------------------------------------------- defaults
def default_context():
    """Context used by the check_* helpers when none is given (test_utils.py:53-62): the thread's default context."""
    return current_context()


def set_default_context(ctx):
    Context.default_ctx = ctx


def default_dtype():
    return np.float32


def get_atol(atol=None):
    return 1e-20 if atol is None else atol


def get_rtol(rtol=None):
    return 1e-5 if rtol is None else rtol


def list_gpus():
    return list(range(num_gpus()))


# ----------------------------------------------------------------------------- random inputs
def random_arrays(*shapes):
    """float32 standard-normal numpy arrays, one per shape (a scalar shape () gives a python float array of size 1)."""
    arrays = [np.array(_RNG.randn(), dtype=default_dtype()) if len(s) == 0 else _RNG.randn(*s).astype(default_dtype())
              for s in shapes]
    return arrays[0] if len(arrays) == 1 else arrays


def random_sample(population, k):
    """k distinct elements in random order."""
    population = list(population)
    idx = _RNG.permutation(len(population))[:k]
    return [population[i] for i in idx]


def rand_shape_2d(dim0=10, dim1=10):
    return _RNG.randint(1, dim0 + 1), _RNG.randint(1, dim1 + 1)


def rand_shape_3d(dim0=10, dim1=10, dim2=10):
    return _RNG.randint(1, dim0 + 1), _RNG.randint(1, dim1 + 1), _RNG.randint(1, dim2 + 1)


def rand_shape_nd(num_dim, dim=10):
    return tuple(_RNG.randint(1, dim + 1, size=num_dim))


def assign_each(the_input, function):
    """Elementwise map of a numpy array through a python function (None: copy)."""
    if function is None:
        return np.array(the_input)
    return np.vectorize(function, otypes=[the_input.dtype])(the_input)


def assign_each2(input1, input2, function):
    if function is None:
        return np.array(input1)
    return np.vectorize(function, otypes=[input1.dtype])(input1, input2)


def rand_sparse_ndarray(shape, stype, density=None, dtype=None, distribution=None, data_init=None, rsp_indices=None,
                        modifier_func=None, shuffle_csr_indices=False, ctx=None):
    """Random row_sparse / csr NDArray plus its numpy components: returns (array, (data, indices[, indptr]))."""
    density = _RNG.rand() if density is None else density
    dtype = default_dtype() if dtype is None else dtype
    if stype == "row_sparse":
        if rsp_indices is not None:
            idx = np.asarray(sorted(rsp_indices), dtype=np.int64)
        else:
            idx = np.nonzero(_RNG.rand(shape[0]) < density)[0].astype(np.int64)
        val = _RNG.rand(len(idx), *shape[1:]).astype(dtype)
        if data_init is not None:
            val.fill(data_init)
        if modifier_func is not None:
            val = assign_each(val, modifier_func)
        arr = nd.sparse.row_sparse_array((val, idx), shape=shape, dtype=dtype) if len(idx) else nd.sparse.zeros("row_sparse", shape, dtype=dtype)
        return arr, (val, idx)
    if stype == "csr":
        assert len(shape) == 2
        dense = _RNG.rand(*shape).astype(dtype)
        if distribution == "powerlaw":
            # row r holds 2^r leading non-zeros until the total budget density * size is used up (the last row takes the rest)
            keep = np.zeros(shape, bool)
            n, left = 1, int(density * shape[0] * shape[1])
            for r in range(shape[0]):
                take = min(shape[1], n, left)
                keep[r, :take] = True
                left -= take
                n *= 2
                if left <= 0:
                    break
        else:
            keep = _RNG.rand(*shape) < density
        dense = dense * keep
        if data_init is not None:
            dense[keep] = data_init
        if modifier_func is not None:
            dense = np.where(keep, assign_each(dense, modifier_func), 0).astype(dtype)
        arr = nd.array(dense, dtype=dtype).tostype("csr")
        return arr, (arr.data.asnumpy(), arr.indices.asnumpy(), arr.indptr.asnumpy())
    raise ValueError("unknown storage type " + str(stype))


def rand_ndarray(shape, stype="default", density=None, dtype=None, modifier_func=None, shuffle_csr_indices=False,
                 distribution=None, ctx=None):
    if stype == "default":
        return nd.array(random_arrays(shape), dtype=dtype, ctx=ctx)
    return rand_sparse_ndarray(shape, stype, density=density, dtype=dtype, modifier_func=modifier_func,
                               distribution=distribution)[0]


def create_sparse_array(shape, stype, data_init=None, rsp_indices=None, dtype=None, modifier_func=None, density=.5,
                        shuffle_csr_indices=False):
    if stype == "row_sparse":
        return rand_sparse_ndarray(shape, stype, density=density, data_init=data_init, rsp_indices=rsp_indices, dtype=dtype,
                                   modifier_func=modifier_func)[0]
    return rand_sparse_ndarray(shape, stype, density=density, data_init=data_init, dtype=dtype, modifier_func=modifier_func)[0]


def create_sparse_array_zd(shape, stype, density, data_init=None, rsp_indices=None, dtype=None, modifier_func=None,
                           shuffle_csr_indices=False):
    """Like create_sparse_array but density 0 yields a well-formed all-zero array."""
    if stype == "row_sparse" and density == 0.0:
        rsp_indices = []
    return create_sparse_array(shape, stype, data_init=data_init, rsp_indices=rsp_indices, dtype=dtype,
                               modifier_func=modifier_func, density=density)


def np_reduce(dat, axis, keepdims, numpy_reduce_func):
    """numpy reduction over several axes with MXNet's keepdims convention."""
    if isinstance(axis, int):
        axis = [axis]
    axis = list(axis) if axis is not None else list(range(dat.ndim))
    ret = dat
    for i in reversed(sorted(a % dat.ndim for a in axis)):
        ret = numpy_reduce_func(ret, axis=i)
    if keepdims:
        shape = list(dat.shape)
        for i in axis:
            shape[i % dat.ndim] = 1
        ret = ret.reshape(tuple(shape))
    return ret


# ----------------------------------------------------------------------------- comparisons
def _np(a):
    return a.asnumpy() if isinstance(a, NDArray) else np.asarray(a)


def same(a, b):
    return np.array_equal(_np(a), _np(b))


def same_array(array1, array2):
    """True when two NDArrays share memory: a write through one is visible through the other."""
    array1[:] += 1
    shared = same(array1, array2)
    array1[:] -= 1
    return shared and same(array1, array2)


def find_max_violation(a, b, rtol=None, atol=None):
    """Index and value of the worst |a-b| / (atol + rtol|b|)."""
    a, b = _np(a), _np(b)
    viol = np.abs(a - b) / (get_atol(atol) + get_rtol(rtol) * np.abs(b))
    loc = np.unravel_index(np.argmax(viol), viol.shape) if viol.ndim else ()
    return loc, float(np.max(viol))


def almost_equal(a, b, rtol=None, atol=None, equal_nan=False):
    return np.allclose(_np(a), _np(b), rtol=get_rtol(rtol), atol=get_atol(atol), equal_nan=equal_nan)


def assert_almost_equal(a, b, rtol=None, atol=None, names=("a", "b"), equal_nan=False):
    a, b = _np(a), _np(b)
    if almost_equal(a, b, rtol, atol, equal_nan=equal_nan):
        return
    loc, viol = find_max_violation(a, b, rtol, atol)
    raise AssertionError("Items are not equal:\nError %f exceeds tolerance rtol=%e, atol=%e.  Location of maximum error: %s, "
                         "%s=%s, %s=%s" % (viol, get_rtol(rtol), get_atol(atol), str(loc), names[0],
                                           str(a[loc] if a.ndim else a), names[1], str(b[loc] if b.ndim else b)))


def almost_equal_ignore_nan(a, b, rtol=None, atol=None):
    a, b = np.atleast_1d(np.copy(_np(a))), np.atleast_1d(np.copy(_np(b)))
    nan = np.logical_or(np.isnan(a), np.isnan(b))
    a[nan] = 0
    b[nan] = 0
    return almost_equal(a, b, rtol, atol)


def assert_almost_equal_ignore_nan(a, b, rtol=None, atol=None, names=("a", "b")):
    a, b = np.copy(_np(a)), np.copy(_np(b))
    nan = np.logical_or(np.isnan(a), np.isnan(b))
    a[nan] = 0
    b[nan] = 0
    assert_almost_equal(a, b, rtol, atol, names)


def assert_exception(f, exception_type, *args, **kwargs):
    try:
        f(*args, **kwargs)
    except exception_type:
        return
    raise AssertionError("%s did not raise %s" % (getattr(f, "__name__", f), exception_type.__name__))


def retry(n):
    """Decorator: re-run a flaky (randomised) test up to n times, pass if any run passes."""
    assert n > 0

    def deco(f):
        @functools.wraps(f)
        def wrapper(*args, **kwargs):
            err = None
            for _ in range(n):
                try:
                    return f(*args, **kwargs)
                except AssertionError as e:
                    err = e
            raise err
        return wrapper
    return deco


def compare_ndarray_tuple(t1, t2, rtol=None, atol=None):
    if t1 is None or t2 is None:
        return
    if isinstance(t1, tuple):
        for s1, s2 in zip(t1, t2):
            compare_ndarray_tuple(s1, s2, rtol, atol)
    else:
        assert_almost_equal(t1, t2, rtol=rtol, atol=atol)


# ----------------------------------------------------------------------------- symbolic helpers
def _is_symbol(x):
    from .symbol import Symbol
    return isinstance(x, Symbol)


def _parse_location(sym, location, ctx, dtype=np.float32):
    """dict name -> NDArray from a list (argument order) or dict of numpy / NDArray values."""
    names = sym.list_arguments()
    if isinstance(location, dict):
        if set(location.keys()) != set(names):
            raise ValueError("Symbol arguments and keys of the given location do not match. symbol args:%s, location.keys():%s"
                             % (str(set(names)), str(set(location.keys()))))
    else:
        location = dict(zip(names, location))
    return {k: (nd.array(v, ctx=ctx, dtype=v.dtype if isinstance(v, np.ndarray) and dtype == "asnumpy" else dtype)
                if isinstance(v, np.ndarray) else (v.as_in_context(ctx) if isinstance(v, NDArray) else nd.array(np.asarray(v), ctx=ctx, dtype=dtype)))
            for k, v in location.items()}


def _parse_aux_states(sym, aux_states, ctx, dtype=np.float32):
    if aux_states is None:
        return None
    names = sym.list_auxiliary_states()
    if not isinstance(aux_states, dict):
        aux_states = dict(zip(names, aux_states))
    elif set(aux_states.keys()) != set(names):
        raise ValueError("Symbol aux_states names and given aux_states do not match.")
    return {k: nd.array(_np(v), ctx=ctx, dtype=dtype) for k, v in aux_states.items()}


def simple_forward(sym, ctx=None, is_train=False, **inputs):
    """Bind numpy inputs, run forward, return numpy output(s)."""
    ctx = ctx or default_context()
    args = {k: nd.array(v, ctx=ctx) for k, v in inputs.items()}
    exe = sym.bind(ctx, args=args, grad_req="null")
    exe.forward(is_train=is_train)
    outs = [o.asnumpy() for o in exe.outputs]
    return outs[0] if len(outs) == 1 else outs


def numeric_grad(executor, location, aux_states=None, eps=1e-4, use_forward_train=True, dtype=np.float32):
    """Central finite differences of sum(outputs[0]) w.r.t. every array in ``location`` (dict name -> numpy).  The
    difference is taken element-wise on the output BEFORE the sum (test_utils.py:numeric_grad of the reference): only the
    elements that depend on the perturbed input contribute, so a float32 executor is not drowned in cancellation noise."""
    for k, v in location.items():
        executor.arg_dict[k][:] = v
    grads = {k: np.zeros(v.shape, dtype=np.float64) for k, v in location.items()}

    def f():
        if aux_states is not None:
            for k, v in aux_states.items():
                executor.aux_dict[k][:] = v
        executor.forward(is_train=use_forward_train)
        return executor.outputs[0].asnumpy().astype(np.float64)
    for k, v in location.items():
        if not np.issubdtype(np.asarray(v).dtype, np.floating):
            continue
        base = np.array(v, dtype=np.float64)
        flat = base.reshape(-1)
        g = grads[k].reshape(-1)
        for i in range(flat.size):
            old = flat[i]
            flat[i] = old + eps / 2
            executor.arg_dict[k][:] = base.reshape(v.shape)
            fp = f()
            flat[i] = old - eps / 2
            executor.arg_dict[k][:] = base.reshape(v.shape)
            fm = f()
            flat[i] = old
            g[i] = float((fp - fm).sum()) / eps
        executor.arg_dict[k][:] = base.reshape(v.shape)
    return grads


def _check_numeric_gradient_fn(fn, inputs, eps=1e-3, rtol=1e-2, atol=1e-3):
    from . import autograd
    arrs = [nd.array(x, dtype="float64") if not isinstance(x, NDArray) else x.astype("float64") for x in inputs]
    for a in arrs:
        a.attach_grad()
    with autograd.record():
        out = fn(*arrs)
    out.backward()
    for a in arrs:
        num = np.zeros(a.shape)
        flat = a._t.detach().reshape(-1)
        for i in range(flat.numel()):
            old = flat[i].item()
            flat[i] = old + eps
            fp = float(fn(*arrs).asnumpy().sum())
            flat[i] = old - eps
            fm = float(fn(*arrs).asnumpy().sum())
            flat[i] = old
            num.reshape(-1)[i] = (fp - fm) / (2 * eps)
        np.testing.assert_allclose(a.grad.asnumpy(), num, rtol=rtol, atol=atol)


def check_numeric_gradient(sym, location, aux_states=None, numeric_eps=1e-3, rtol=1e-2, atol=None, grad_nodes=None,
                           use_forward_train=True, ctx=None, grad_stype_dict=None, dtype=np.float64, eps=None):
    """Symbolic gradients of ``sum(sym * random_projection)`` against central finite differences
    (test_utils.py:790-924).  ``sym`` may also be a python function of NDArrays returning a scalar NDArray, with
    ``location`` the list of its inputs (this repo's short form)."""
    if not _is_symbol(sym):
        return _check_numeric_gradient_fn(sym, location, eps=eps or numeric_eps, rtol=rtol, atol=1e-3 if atol is None else atol)
    from . import symbol as S
    ctx = ctx or default_context()
    atol = 1e-3 if atol is None else atol
    loc = _parse_location(sym, location, ctx, dtype=dtype)
    loc_np = {k: v.asnumpy() for k, v in loc.items()}
    aux = _parse_aux_states(sym, aux_states, ctx, dtype=dtype)
    aux_np = {k: v.asnumpy() for k, v in aux.items()} if aux is not None else None
    if grad_nodes is None:
        grad_nodes = {k: "write" for k in sym.list_arguments()}
    elif isinstance(grad_nodes, (list, tuple)):
        grad_nodes = {k: "write" for k in grad_nodes}
    out_shape = sym.infer_shape(**{k: v.shape for k, v in loc.items()})[1][0]
    proj = S.var("__random_proj")
    loss = S.make_loss(sym[0] * proj if len(sym.list_outputs()) > 1 else sym * proj)
    args = dict(loc)
    args["__random_proj"] = nd.array(_RNG.uniform(0.5, 1.5, out_shape), ctx=ctx, dtype=dtype)
    req = {k: grad_nodes.get(k, "null") for k in sym.list_arguments()}
    req["__random_proj"] = "null"
    grads0 = {k: nd.array(_RNG.normal(size=loc[k].shape), ctx=ctx, dtype=dtype) for k in grad_nodes}
    bound_grads = {k: v.copy() for k, v in grads0.items()}
    exe = loss.bind(ctx, args=args, args_grad=bound_grads, grad_req=req, aux_states=aux)
    exe.forward(is_train=True)
    exe.backward()
    sym_grads = {k: (exe.grad_dict[k] if k in exe.grad_dict else bound_grads[k]).asnumpy() for k in grad_nodes}
    num_loc = dict(loc_np)
    num_loc["__random_proj"] = args["__random_proj"].asnumpy()
    fd = numeric_grad(exe, num_loc, aux_np, eps=numeric_eps, use_forward_train=use_forward_train, dtype=dtype)
    for name, kind in grad_nodes.items():
        expect = fd[name] + (grads0[name].asnumpy() if kind == "add" else 0)
        if kind == "null":
            expect = grads0[name].asnumpy()
        assert_almost_equal(expect, sym_grads[name], rtol, atol, ("NUMERICAL_%s" % name, "BACKWARD_%s" % name))


def check_symbolic_forward(sym, location, expected, rtol=1E-4, atol=None, aux_states=None, ctx=None, equal_nan=False,
                           dtype=np.float32):
    """Outputs of ``sym`` on ``location`` vs numpy ``expected`` (list in output order or dict by output name)."""
    ctx = ctx or default_context()
    loc = _parse_location(sym, location, ctx, dtype=dtype)
    aux = _parse_aux_states(sym, aux_states, ctx, dtype=dtype)
    if isinstance(expected, dict):
        expected = [expected[k] for k in sym.list_outputs()]
    exe = sym.bind(ctx, args=loc, aux_states=aux, grad_req="null")
    exe.forward(is_train=False)
    for name, exp, out in zip(sym.list_outputs(), expected, exe.outputs):
        assert_almost_equal(exp, out.asnumpy(), rtol, 1e-20 if atol is None else atol, ("EXPECTED_%s" % name, "FORWARD_%s" % name),
                            equal_nan=equal_nan)
    return exe.outputs


def check_symbolic_backward(sym, location, out_grads, expected, rtol=1e-5, atol=None, aux_states=None, grad_req="write",
                            ctx=None, grad_stypes=None, equal_nan=False, dtype=np.float32):
    """Argument gradients for given head gradients vs numpy ``expected`` (list in argument order or dict)."""
    ctx = ctx or default_context()
    loc = _parse_location(sym, location, ctx, dtype=dtype)
    aux = _parse_aux_states(sym, aux_states, ctx, dtype=dtype)
    names = sym.list_arguments()
    if not isinstance(expected, dict):
        expected = dict(zip(names, expected))
    if isinstance(grad_req, str):
        grad_req = {k: grad_req for k in names}
    elif isinstance(grad_req, (list, tuple)):
        grad_req = dict(zip(names, grad_req))
    init = {k: _RNG.normal(size=loc[k].shape).astype(np.float64) for k in expected}
    args_grad = {k: nd.array(v, ctx=ctx, dtype=dtype) for k, v in init.items()}
    exe = sym.bind(ctx, args=loc, args_grad=args_grad, aux_states=aux,
                   grad_req={k: (grad_req.get(k, "null") if k in expected else "null") for k in names})
    exe.forward(is_train=True)
    if isinstance(out_grads, dict):
        out_grads = [out_grads[k] for k in sym.list_outputs()]
    if out_grads is not None:
        out_grads = [g if isinstance(g, NDArray) else nd.array(g, ctx=ctx, dtype=dtype) for g in out_grads]
    exe.backward(out_grads)
    got_nd = {k: (exe.grad_dict[k] if k in exe.grad_dict else args_grad[k]) for k in expected}
    if grad_stypes:
        got_nd = {k: (v.tostype(grad_stypes[k]) if k in grad_stypes else v) for k, v in got_nd.items()}
    got = {k: v.asnumpy() for k, v in got_nd.items()}
    for name in expected:
        req = grad_req.get(name, "null")
        if req == "write":
            assert_almost_equal(expected[name], got[name], rtol, 1e-20 if atol is None else atol,
                                ("EXPECTED_%s" % name, "BACKWARD_%s" % name), equal_nan=equal_nan)
        elif req == "add":
            assert_almost_equal(_np(expected[name]) + init[name].astype(got[name].dtype), got[name], rtol,
                                1e-20 if atol is None else atol, ("EXPECTED_%s" % name, "BACKWARD_%s" % name), equal_nan=equal_nan)
        elif req == "null":
            assert_almost_equal(init[name].astype(got[name].dtype), got[name], rtol, 1e-20 if atol is None else atol,
                                ("EXPECTED_%s" % name, "BACKWARD_%s" % name), equal_nan=equal_nan)
        else:
            raise ValueError("Invalid grad_req %s for argument %s" % (req, name))
    return got_nd             # NDArrays (with the requested gradient storage types), like the reference


def check_speed(sym, location=None, ctx=None, N=20, grad_req=None, typ="whole", **kwargs):
    """Seconds per iteration of forward+backward (typ='whole') or forward only (typ='forward')."""
    ctx = ctx or default_context()
    grad_req = grad_req or "write"
    if location is None:
        exe = sym.simple_bind(grad_req=grad_req, ctx=ctx, **kwargs)
        location = {k: _RNG.normal(size=arr.shape, scale=1.0) for k, arr in exe.arg_dict.items()}
    else:
        exe = sym.simple_bind(grad_req=grad_req, ctx=ctx, **{k: v.shape for k, v in location.items()})
    for name, iarr in location.items():
        exe.arg_dict[name][:] = nd.array(np.asarray(iarr), ctx=ctx).astype(exe.arg_dict[name].dtype)

    def sync():
        nd.waitall()
        if torch.cuda.is_available():
            torch.cuda.synchronize()
    if typ not in ("whole", "forward"):
        raise ValueError('typ can only be "whole" or "forward".')

    def step():
        exe.forward(is_train=(typ == "whole"))
        if typ == "whole":
            exe.backward(out_grads=[nd.ones(o.shape, ctx=ctx, dtype=o.dtype) for o in exe.outputs])
    step()
    sync()
    tic = time.time()
    for _ in range(N):
        step()
    sync()
    return (time.time() - tic) / N


def _check_consistency_fn(fn, inputs, ctx_list, dtypes=("float32",), rtol=1e-3, atol=1e-4):
    ref = None
    for ctx in ctx_list:
        for dt in dtypes:
            out = fn(*[nd.array(x, ctx=ctx, dtype=dt) for x in inputs]).asnumpy().astype(np.float64)
            if ref is None:
                ref = out
            else:
                np.testing.assert_allclose(out, ref, rtol=rtol, atol=atol)


def check_consistency(sym, ctx_list, *args, **kwargs):
    """The same symbol bound on several (context, dtype) combinations must produce the same outputs and gradients
    (test_utils.py:1207-1366).  ``ctx_list``: list of dicts ``{'ctx': Context, <input name>: shape, 'type_dict': {...}}``;
    the run with the widest dtype is the reference.  Keyword arguments as in the reference: scale, grad_req, arg_params,
    aux_params, tol, raise_on_err, ground_truth, equal_nan, use_uniform.
    Short form for plain functions: ``check_consistency(fn, inputs, [ctx, ...], dtypes=(...), rtol=, atol=)``."""
    if callable(sym) and not _is_symbol(sym):
        ctxs = args[0] if args else kwargs.pop("ctxs", [default_context()])
        return _check_consistency_fn(sym, ctx_list, ctxs, **kwargs)
    return _check_consistency_sym(sym, ctx_list, *args, **kwargs)


def _check_consistency_sym(sym, ctx_list, scale=1.0, grad_req="write", arg_params=None, aux_params=None, tol=None,
                           raise_on_err=True, ground_truth=None, equal_nan=False, use_uniform=False):
    width = {"float16": 0, "bfloat16": 0, "float32": 1, "float64": 2, "uint8": -1, "int32": -1, "int64": -1}
    if tol is None:
        tol = {"float16": 1e-1, "bfloat16": 1e-1, "float32": 1e-3, "float64": 1e-5, "uint8": 0, "int32": 0, "int64": 0}
    elif isinstance(tol, numbers.Number):
        tol = {k: tol for k in width}
    else:
        tol = {np.dtype(k).name if not isinstance(k, str) else k: v for k, v in tol.items()}
    syms = list(sym) if isinstance(sym, (list, tuple)) else [sym] * len(ctx_list)
    assert len(syms) == len(ctx_list) >= 2
    out_names = syms[0].list_outputs()
    arg_names = syms[0].list_arguments()
    exes = []
    for s, spec in zip(syms, ctx_list):
        spec = dict(spec)
        ctx = spec.pop("ctx")
        type_dict = spec.pop("type_dict", None)
        exes.append(s.simple_bind(ctx, grad_req=grad_req, type_dict=type_dict, **spec))
    arg_params = dict(arg_params or {})
    aux_params = dict(aux_params or {})
    for n, arr in exes[0].arg_dict.items():
        if n not in arg_params:
            arg_params[n] = (_RNG.uniform(-0.92 * scale, 0.92 * scale, arr.shape) if use_uniform
                             else _RNG.normal(size=arr.shape, scale=scale))
    for n in exes[0].aux_dict:
        aux_params.setdefault(n, 0)
    for exe in exes:
        for n, arr in exe.arg_dict.items():
            arr[:] = nd.array(np.asarray(arg_params[n]), ctx=arr.context).astype(arr.dtype)
        for n, arr in exe.aux_dict.items():
            arr[:] = aux_params[n]

    def dname(a):
        from .base import dtype_name
        return dtype_name(a._t.dtype)
    dts = [dname(exe.outputs[0]) if exe.outputs else dname(list(exe.arg_dict.values())[0]) for exe in exes]
    for exe in exes:
        exe.forward(is_train=False)
    dts = [dname(exe.outputs[0]) for exe in exes]
    ref_i = int(np.argmax([width.get(d, 1) for d in dts]))
    gt = ground_truth
    if gt is None:
        gt = {n: v.asnumpy() for n, v in exes[ref_i].output_dict.items()}
    errors = []

    def compare(stage):
        for i, exe in enumerate(exes):
            if i == ref_i and ground_truth is None:
                continue
            t = tol[dts[i]]
            got = dict(exe.output_dict)
            if stage == "backward":
                got.update({k: v for k, v in exe.grad_dict.items() if v is not None})
            for name, arr in got.items():
                if name not in gt:
                    continue
                try:
                    assert_almost_equal(arr.asnumpy().astype(np.float64), np.asarray(gt[name], dtype=np.float64), rtol=t, atol=t,
                                        equal_nan=equal_nan)
                except AssertionError as e:
                    errors.append("%s check failed for context %d (%s) on %s: %s" % (stage, i, dts[i], name, e))
    compare("forward")
    if grad_req != "null":
        for exe in exes:
            exe.forward(is_train=True)
            exe.backward([nd.array(np.asarray(gt[n]), ctx=o.context).astype(o.dtype) for n, o in zip(out_names, exe.outputs)]
                         if ground_truth is None else exe.outputs)
        if ground_truth is None:
            gt = {n: v.asnumpy() for n, v in exes[ref_i].output_dict.items()}
            gt.update({n: v.asnumpy() for n, v in exes[ref_i].grad_dict.items() if v is not None})
        compare("backward")
    if errors and raise_on_err:
        raise AssertionError("\n".join(errors))
    return gt


# ----------------------------------------------------------------------------- data sets (no network in this image)
def download(url, fname=None, dirname=None, overwrite=False, retries=5):
    """Return the local path if the file already exists; fetching needs network access, which a sandboxed host lacks."""
    fname = fname or url.split("/")[-1]
    if dirname is not None:
        os.makedirs(dirname, exist_ok=True)
        fname = os.path.join(dirname, fname)
    if os.path.exists(fname) and not overwrite:
        return fname
    try:
        import urllib.request
        for _ in range(max(1, retries)):
            try:
                urllib.request.urlretrieve(url, fname)
                return fname
            except Exception as e:  # noqa: BLE001
                err = e
        raise err
    except Exception as e:  # noqa: BLE001
        raise MXNetError("download(%s) failed: %s" % (url, e))


def get_mnist(path="data"):
    """MNIST as numpy arrays from ``path`` (train-images-idx3-ubyte.gz etc. must already be there)."""
    import gzip
    import struct

    def read(label_f, image_f):
        with gzip.open(os.path.join(path, label_f)) as f:
            struct.unpack(">II", f.read(8))
            label = np.frombuffer(f.read(), dtype=np.int8).astype(np.int32)
        with gzip.open(os.path.join(path, image_f), "rb") as f:
            _, _, rows, cols = struct.unpack(">IIII", f.read(16))
            image = np.frombuffer(f.read(), dtype=np.uint8).reshape(len(label), 1, rows, cols).astype(np.float32) / 255
        return label, image
    for f in ("train-labels-idx1-ubyte.gz", "train-images-idx3-ubyte.gz", "t10k-labels-idx1-ubyte.gz", "t10k-images-idx3-ubyte.gz"):
        if not os.path.exists(os.path.join(path, f)):
            raise MXNetError("MNIST file %s not found under %s (no network access to fetch it)" % (f, path))
    tl, ti = read("train-labels-idx1-ubyte.gz", "train-images-idx3-ubyte.gz")
    vl, vi = read("t10k-labels-idx1-ubyte.gz", "t10k-images-idx3-ubyte.gz")
    return {"train_data": ti, "train_label": tl, "test_data": vi, "test_label": vl}


def _needs_network(what):
    raise MXNetError("%s downloads its data set; this host has no network access. Place the files manually." % what)


def get_mnist_pkl():
    if not os.path.isfile("data/mnist.pkl.gz"):
        _needs_network("get_mnist_pkl")


def get_mnist_ubyte():
    """MNIST idx files under ./data (train-images-idx3-ubyte ...), as the reference's C++ MNISTIter expects them."""
    need = ["data/train-images-idx3-ubyte", "data/train-labels-idx1-ubyte", "data/t10k-images-idx3-ubyte", "data/t10k-labels-idx1-ubyte"]
    if not all(os.path.isfile(f) for f in need):
        _needs_network("get_mnist_ubyte")


def get_cifar10():
    """CIFAR-10 RecordIO files under ./data/cifar."""
    need = ["data/cifar/train.rec", "data/cifar/test.rec", "data/cifar/train.lst", "data/cifar/test.lst"]
    if not all(os.path.isfile(f) for f in need):
        _needs_network("get_cifar10")


def get_zip_data(data_dir, url, data_origin_name):
    if not os.path.isfile(os.path.join(data_dir, data_origin_name)):
        _needs_network("get_zip_data")


def get_bz2_data(data_dir, data_name, url, data_origin_name):
    if not os.path.isfile(os.path.join(data_dir, data_name)):
        _needs_network("get_bz2_data")


def get_im2rec_path(home_env="MXNET_HOME"):
    """Path of the im2rec tool (python version shipped in tools/)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return os.path.join(here, "tools", "im2rec.py")


def get_mnist_iterator(batch_size, input_shape, num_parts=1, part_index=0, path="data"):
    from . import io
    m = get_mnist(path)
    flat = len(input_shape) == 1

    def shard(a):
        n = len(a) // num_parts
        return a[part_index * n:(part_index + 1) * n]
    tr, te = m["train_data"], m["test_data"]
    if flat:
        tr, te = tr.reshape(len(tr), -1), te.reshape(len(te), -1)
    return (io.NDArrayIter(shard(tr), shard(m["train_label"]), batch_size, shuffle=True),
            io.NDArrayIter(te, m["test_label"], batch_size))


# ----------------------------------------------------------------------------- environment / misc
def set_env_var(key, val, default_val=""):
    """Set an environment variable, return its previous value (or ``default_val``)."""
    prev = os.environ.get(key, default_val)
    os.environ[key] = val
    return prev


@contextlib.contextmanager
def discard_stderr():
    """Silence writes to file descriptor 2 inside the block (native code included)."""
    try:
        fd = sys.stderr.fileno()
    except Exception:  # noqa: BLE001 -- captured stderr (pytest): nothing to redirect
        yield
        return
    saved = os.dup(fd)
    with open(os.devnull, "w") as sink:
        try:
            os.dup2(sink.fileno(), fd)
            yield
        finally:
            os.dup2(saved, fd)
            os.close(saved)


class DummyIter(object):
    """Repeats the first batch of a real iterator forever: removes IO from speed measurements."""

    def __init__(self, real_iter):
        self.real_iter = real_iter
        self.provide_data = real_iter.provide_data
        self.provide_label = real_iter.provide_label
        self.batch_size = real_iter.batch_size
        self.the_batch = next(iter(real_iter))

    def __iter__(self):
        return self

    def next(self):
        return self.the_batch

    __next__ = next

    def reset(self):
        pass


# ----------------------------------------------------------------------------- statistical checks of samplers
def gen_buckets_probs_with_ppf(ppf, nbuckets):
    """Equal-probability buckets of a continuous distribution from its percent-point function."""
    assert nbuckets > 0
    p = 1.0 / nbuckets
    buckets = [(ppf(i * p), ppf((i + 1) * p)) for i in range(nbuckets)]
    return buckets, [p] * nbuckets


def mean_check(generator, mu, sigma, nsamples=1000000):
    """Sample mean within 3 standard errors of mu."""
    s = np.asarray(generator(nsamples), dtype=np.float64)
    return abs(s.mean() - mu) < 3 * sigma / np.sqrt(nsamples)


def var_check(generator, sigma, nsamples=1000000):
    """Sample variance within 3 standard errors of sigma^2 (normal approximation)."""
    s = np.asarray(generator(nsamples), dtype=np.float64)
    return abs(s.var() - sigma ** 2) < 3 * np.sqrt(2 * sigma ** 4 / (nsamples - 1))


def chi_square_check(generator, buckets, probs, nsamples=1000000):
    """Pearson chi-square test of ``generator(nsamples)`` against bucket probabilities.  Buckets are (lo, hi) intervals for
    continuous generators or plain values for discrete ones.  Returns (p value, observed counts, expected counts)."""
    s = np.asarray(generator(nsamples)).reshape(-1)
    expected = np.asarray(probs, dtype=np.float64) * len(s)
    if isinstance(buckets[0], (tuple, list)):
        obs = np.array([np.logical_and(s >= lo, s < hi).sum() for lo, hi in buckets], dtype=np.float64)
    else:
        obs = np.array([(s == b).sum() for b in buckets], dtype=np.float64)
    import scipy.stats as ss
    stat = ((obs - expected) ** 2 / np.maximum(expected, 1e-12)).sum()
    p = float(ss.chi2.sf(stat, len(buckets) - 1))
    return p, obs, expected


def verify_generator(generator, buckets, probs, nsamples=1000000, nrepeat=5, success_rate=0.15, alpha=0.05):
    """Repeat the chi-square test; at least ``success_rate`` of the runs must not reject at level alpha."""
    ps = [chi_square_check(generator, buckets, probs, nsamples)[0] for _ in range(nrepeat)]
    ok = sum(p > alpha for p in ps)
    if ok < nrepeat * success_rate:
        raise AssertionError("Generator test fails, Chi-square p=%s, buckets=%s, probs=%s" % (ps, buckets, probs))
    return ps


def compare_optimizer(opt1, opt2, shape, dtype, w_stype="default", g_stype="default", rtol=1e-4, atol=1e-5,
                      compare_states=True):
    """One random weight / gradient through both optimizers (update_multi_precision); weights and states must agree."""
    if w_stype == "default":
        w2 = rand_ndarray(shape, dtype=dtype)
        w1 = w2.copy()
    else:
        w2 = rand_ndarray(shape, w_stype, density=1, dtype=dtype)
        w1 = w2.copy().tostype("default")
    if g_stype == "default":
        g2 = rand_ndarray(shape, dtype=dtype)
        g1 = g2.copy()
    else:
        g2 = rand_ndarray(shape, g_stype, dtype=dtype)
        g1 = g2.copy().tostype("default")
    s1 = opt1.create_state_multi_precision(0, w1)
    s2 = opt2.create_state_multi_precision(0, w2)
    if compare_states:
        compare_ndarray_tuple(s1, s2)
    opt1.update_multi_precision(0, w1, g1, s1)
    opt2.update_multi_precision(0, w2, g2, s2)
    if compare_states:
        compare_ndarray_tuple(s1, s2, rtol=rtol, atol=atol)
    assert_almost_equal(w1.asnumpy(), (w2.tostype("default") if w_stype != "default" else w2).asnumpy(), rtol=rtol, atol=atol)
