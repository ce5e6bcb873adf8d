"""Custom operators written in python (reference: python/mxnet/operator.py CustomOp/CustomOpProp,
src/operator/custom/custom.cc).  A registered CustomOp becomes callable as ``mx.nd.Custom(..., op_type=name)``
and differentiable through torch autograd."""
import torch

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

    def list_outputs(self):
        return ["output"]

    def list_arguments(self):
        return ["data"]

    def list_auxiliary_states(self):
        return []

    def create_operator(self, ctx, in_shapes, in_dtypes):
        return CustomOp()


def register(reg_name):
    def do_register(prop_cls):
        _REG[reg_name] = prop_cls
        return prop_cls
    return do_register


def Custom(*inputs, op_type=None, **kwargs):
    prop = _REG[op_type](**{k: str(v) for k, v in kwargs.items()})
    in_shapes = [list(i.shape) for i in inputs]
    _, out_shapes, _ = prop.infer_shape(in_shapes)
    opr = prop.create_operator(inputs[0].context, in_shapes, [i.dtype for i in inputs])
    ctx0 = inputs[0].context
    is_train = autograd.is_training()

    class _Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, *ts):
            ins = [NDArray(t.detach(), ctx0) for t in ts]
            outs = [NDArray(torch.zeros(tuple(s), dtype=ts[0].dtype, device=ts[0].device), ctx0) for s in out_shapes]
            with autograd.pause():
                opr.forward(is_train, ["write"] * len(outs), ins, outs, [])
            ctx.ins, ctx.outs = ins, outs
            return tuple(o._t for o in outs)

        @staticmethod
        def backward(ctx, *gs):
            in_grads = [NDArray(torch.zeros_like(i._t), ctx0) for i in ctx.ins]
            with autograd.pause():
                opr.backward(["write"] * len(in_grads), [NDArray(g, ctx0) for g in gs], ctx.ins, ctx.outs, in_grads, [])
            return tuple(g._t for g in in_grads)

    res = _op(lambda *ts: list(_Fn.apply(*ts)), *inputs)
    return res[0] if len(res) == 1 else res


_ndop._REG["Custom"] = Custom
