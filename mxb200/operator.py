"""Operators written in python (reference: python/mxnet/operator.py ``CustomOp`` / ``CustomOpProp`` / ``register`` and the
legacy ``PythonOp`` / ``NumpyOp`` / ``NDArrayOp``; C++ side src/operator/custom/custom.cc).

A registered ``CustomOpProp`` becomes callable as ``mx.nd.Custom(*inputs, op_type=name, **params)`` and
``mx.sym.Custom(...)``.  Inputs are the prop's ``list_arguments()`` followed by its ``list_auxiliary_states()`` (positional
or by keyword); the operator is differentiable through torch autograd: ``CustomOp.backward`` runs as the backward of a
``torch.autograd.Function`` and receives the output gradients only when the prop asked for them (``need_top_grad``).
Auxiliary states are handed to forward / backward as writable arrays and never receive a gradient.

The reference runs custom operators on a dedicated worker thread so that a python callback cannot block the engine
(custom-inl.h CustomOperator::Push); here the call is made inline on the calling thread -- kernels it launches are
ordered by the CUDA stream like every other operator.
"""
import numpy as np
import torch

from .base import MXNetError, dtype_name, np_dtype
from .ndarray import NDArray, op as _ndop
from .ndarray.ndarray import _op
from . import autograd

_REG = {}


class CustomOp(object):
    def forward(self, is_train, req, in_data, out_data, aux):
        raise NotImplementedError

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        raise NotImplementedError

    def assign(self, dst, req, src):
        if req == "null":
            return
        if req in ("write", "inplace"):
            dst[:] = src
        elif req == "add":
            dst[:] = dst + src


class CustomOpProp(object):
    def __init__(self, need_top_grad=True):
        self.need_top_grad_ = need_top_grad

    def infer_shape(self, in_shape):
        return in_shape, (in_shape[0],) * len(self.list_outputs()), ()

    def infer_type(self, in_type):
        return in_type, [in_type[0]] * len(self.list_outputs()), [in_type[0]] * len(self.list_auxiliary_states())

    def infer_storage_type(self, in_stype):
        return in_stype, ["default"] * len(self.list_outputs()), ["default"] * len(self.list_auxiliary_states())

    def infer_storage_type_backward(self, ograd_stype, in_stype, out_stype, igrad_stype, aux_stype):
        return ograd_stype, in_stype, out_stype, ["default"] * len(igrad_stype), aux_stype

    def list_outputs(self):
        return ["output"]

    def list_arguments(self):
        return ["data"]

    def list_auxiliary_states(self):
        return []

    def declare_backward_dependency(self, out_grad, in_data, out_data):
        deps = []
        if self.need_top_grad_:
            deps.extend(out_grad)
        deps.extend(in_data)
        deps.extend(out_data)
        return deps

    def create_operator(self, ctx, in_shapes, in_dtypes):
        return CustomOp()


def register(reg_name):
    def do_register(prop_cls):
        _REG[reg_name] = prop_cls
        return prop_cls
    return do_register


def get_prop(op_type, params):
    """The CustomOpProp instance of a registered operator; ``params`` reach its constructor as strings (custom.cc)."""
    if op_type not in _REG:
        raise MXNetError("Custom: operator type %r is not registered" % (op_type,))
    return _REG[op_type](**{k: str(v) for k, v in params.items()})


def _prop_params(kwargs):
    return {k: v for k, v in kwargs.items() if not isinstance(v, NDArray) and not (k.startswith("__") and k.endswith("__"))
            and k not in ("name", "out")}


def Custom(*inputs, op_type=None, **kwargs):
    kwargs.pop("name", None)
    out = kwargs.pop("out", None)
    prop = get_prop(op_type, _prop_params(kwargs))
    arg_names, aux_names = list(prop.list_arguments()), list(prop.list_auxiliary_states())
    by_name = {k: v for k, v in kwargs.items() if isinstance(v, NDArray)}
    pos = list(inputs)
    ordered = []
    for n in arg_names + aux_names:
        if n in by_name:
            ordered.append(by_name[n])
        elif pos:
            ordered.append(pos.pop(0))
        else:
            raise MXNetError("Custom(%s): input %r is missing" % (op_type, n))
    args, auxs = ordered[:len(arg_names)], ordered[len(arg_names):]
    in_shapes = [list(i.shape) for i in args]
    _ishape, out_shapes, _aux_shapes = prop.infer_shape(in_shapes)
    in_types = [np_dtype(i._t.dtype) for i in args]
    try:
        _it, out_types, _at = prop.infer_type(in_types)
    except Exception:  # noqa: BLE001 -- props written for a single dtype
        out_types = [in_types[0]] * len(out_shapes)
    from .context import current_context
    ctx0 = args[0].context if args else current_context()
    opr = prop.create_operator(ctx0, in_shapes, in_types)
    is_train = autograd.is_training()
    need_top = bool(getattr(prop, "need_top_grad_", True))
    if not args:                                                # a source operator: nothing to differentiate
        from .base import torch_dtype
        outs = [NDArray(torch.zeros(tuple(s), dtype=torch_dtype(dtype_name(t)), device=ctx0.torch_device), ctx0)
                for s, t in zip(out_shapes, out_types)]
        with autograd.pause(train_mode=is_train):
            opr.forward(is_train, ["write"] * len(outs), [], outs, auxs)
        return outs[0] if len(outs) == 1 else outs
    if args[0]._t.device.type == "meta":                        # shape inference
        from .base import torch_dtype
        outs = [NDArray(torch.empty(tuple(s), dtype=torch_dtype(dtype_name(t)), device="meta"), ctx0) for s, t in zip(out_shapes, out_types)]
        return outs[0] if len(outs) == 1 else outs
    is_sparse = [getattr(a, "stype", "default") != "default" for a in args]
    out_stypes = ["default"] * len(out_shapes)
    if any(is_sparse):
        try:
            out_stypes = list(prop.infer_storage_type([getattr(a, "stype", "default") for a in args])[1])
        except Exception:  # noqa: BLE001 -- a prop without storage inference works on dense outputs
            pass

    class _Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, *ts):
            from .base import torch_dtype
            ins = [a if sp else NDArray(t.detach(), ctx0) for a, t, sp in zip(args, ts, is_sparse)]
            outs = [NDArray(torch.zeros(tuple(s), dtype=torch_dtype(dtype_name(t)), device=ts[0].device), ctx0)
                    for s, t in zip(out_shapes, out_types)]
            if any(st != "default" for st in out_stypes):
                from .ndarray import sparse as _sp
                outs = [o if st == "default" else _sp.zeros(st, o.shape, ctx=ctx0, dtype=dtype_name(o._t.dtype))
                        for o, st in zip(outs, out_stypes)]
            with autograd.pause(train_mode=is_train):
                opr.forward(is_train, ["write"] * len(outs), ins, outs, auxs)
            ctx.ins, ctx.outs = ins, outs
            return tuple((o.tostype("default") if getattr(o, "stype", "default") != "default" else o)._t for o in outs)

        @staticmethod
        def backward(ctx, *gs):
            in_grads = [NDArray(torch.zeros(tuple(i.shape), dtype=i._t.dtype if hasattr(i, "_t") else torch.float32,
                                            device=gs[0].device), ctx0) for i in ctx.ins]
            ograds = [NDArray(g.contiguous(), ctx0) for g in gs] if need_top else \
                [NDArray(torch.zeros_like(g), ctx0) for g in gs]
            with autograd.pause():
                opr.backward(["write"] * len(in_grads), ograds, ctx.ins, ctx.outs, in_grads, auxs)
            return tuple((g.tostype("default") if getattr(g, "stype", "default") != "default" else g)._t for g in in_grads)

    # a sparse argument enters the tape through its dense leaf (the array autograd accumulates its gradient in)
    dense_args = [NDArray(a._t, ctx0) if sp else a for a, sp in zip(args, is_sparse)]
    res = _op(lambda *ts: list(_Fn.apply(*ts)), *dense_args)
    if out is not None:
        for o, r in zip(out if isinstance(out, (list, tuple)) else [out], res):
            o[:] = r
        return out
    return res[0] if len(res) == 1 else res


_ndop._REG["Custom"] = Custom
import sys as _sys  # noqa: E402
setattr(_sys.modules[NDArray.__module__.rsplit(".", 1)[0]], "Custom", Custom)      # mx.nd.Custom


# ---- legacy python operators (operator.py:42 PythonOp, :150 NumpyOp, :253 NDArrayOp) ---------------------------------------
class PythonOp(object):
    """Pre-``CustomOp`` interface: one object is both the property and the operator; ``op(*symbols)`` /
    ``op.get_symbol(...)`` builds the graph node.  Kept for scripts of that era; implemented on the Custom machinery."""

    _counter = [0]

    def __init__(self, need_top_grad=True):
        self.need_top_grad_ = need_top_grad
        self._op_type = None

    def __call__(self, *args, **kwargs):
        return self.get_symbol(*args, **kwargs)

    def _arrays(self, xs):
        return xs

    def get_symbol(self, *args, **kwargs):
        if self._op_type is None:
            PythonOp._counter[0] += 1
            self._op_type = "_legacy_python_op_%d" % PythonOp._counter[0]
            outer = self

            class _Bridge(CustomOp):
                def forward(self, is_train, req, in_data, out_data, aux):
                    outer.forward(in_data=outer._arrays(in_data), out_data=outer._arrays(out_data))
                    outer._writeback(out_data)

                def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
                    outer.backward(out_grad=outer._arrays(out_grad), in_data=outer._arrays(in_data), out_data=outer._arrays(out_data),
                                   in_grad=outer._arrays(in_grad))
                    outer._writeback(in_grad)

            class _Prop(CustomOpProp):
                def __init__(self):
                    super(_Prop, self).__init__(need_top_grad=outer.need_top_grad())

                def list_arguments(self):
                    return outer.list_arguments()

                def list_outputs(self):
                    return outer.list_outputs()

                def infer_shape(self, in_shape):
                    r = outer.infer_shape(in_shape)
                    return r[0], r[1], []

                def create_operator(self, ctx, in_shapes, in_dtypes):
                    return _Bridge()
            _REG[self._op_type] = _Prop
        from . import symbol
        return symbol.Custom(*args, op_type=self._op_type, **kwargs)

    def _writeback(self, arrays):
        pass

    def forward(self, in_data, out_data):
        out_data[0][:] = in_data[0]

    def backward(self, out_grad, in_data, out_data, in_grad):
        in_grad[0][:] = 1.0

    def infer_shape(self, in_shape):
        return in_shape, [in_shape[0]]

    def list_outputs(self):
        return ["output"]

    def list_arguments(self):
        return ["data"]

    def need_top_grad(self):
        return self.need_top_grad_


class NumpyOp(PythonOp):
    """forward / backward see numpy arrays (written in place)."""

    def _arrays(self, xs):
        # CPU arrays are handed out as numpy VIEWS of the same memory; device arrays as host copies written back afterwards
        self._pairs = getattr(self, "_pairs", [])
        out = []
        for x in xs:
            t = x._t.detach()
            if t.device.type == "cpu" and t.is_contiguous() and t.dtype != torch.bfloat16:
                out.append(t.numpy())
            else:
                h = t.float().cpu().numpy().copy()
                self._pairs.append((x, h))
                out.append(h)
        return out

    def _writeback(self, arrays):
        for x, h in getattr(self, "_pairs", []):
            if any(x is a for a in arrays):
                x[:] = h
        self._pairs = []


class NDArrayOp(PythonOp):
    """forward / backward see NDArrays."""

    def declare_backward_dependency(self, out_grad, in_data, out_data):
        deps = []
        if self.need_top_grad():
            deps.extend(out_grad)
        deps.extend(in_data)
        deps.extend(out_data)
        return deps
