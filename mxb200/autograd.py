"""Autograd scope API on top of torch.autograd.

Reference: python/mxnet/autograd.py, src/imperative/imperative.cc (tape
recording, ``Imperative::Backward``).  torch's tape replaces the reference's
NNVM tape; this module reproduces the *scoping* semantics: ops are taped only
inside ``record()``, ``train_mode``/``predict_mode`` toggle layer behaviour,
and ``grad_req='write'`` overwrites gradients on each backward while ``'add'``
accumulates.
"""
import threading
import weakref

import torch

__all__ = ["record", "pause", "train_mode", "predict_mode", "is_recording", "is_training",
           "backward", "grad", "mark_variables", "set_recording", "set_training", "Function"]

_state = threading.local()
_bw_epoch = [0]


def _st():
    if not hasattr(_state, "rec"):
        _state.rec = False
        _state.train = False
    return _state


def is_recording():
    return _st().rec


def is_training():
    return _st().train


def set_recording(flag):
    s = _st()
    old, s.rec = s.rec, bool(flag)
    return old


def set_training(flag):
    s = _st()
    old, s.train = s.train, bool(flag)
    return old


class _Scope(object):
    def __init__(self, rec, train):
        self._rec, self._train = rec, train
        self._old = None
        self._grad_ctx = None

    def __enter__(self):
        s = _st()
        self._old = (s.rec, s.train)
        if self._rec is not None:
            s.rec = self._rec
            self._grad_ctx = torch.enable_grad() if self._rec else torch.no_grad()
            self._grad_ctx.__enter__()
        if self._train is not None:
            s.train = self._train
        return self

    def __exit__(self, *a):
        s = _st()
        if self._grad_ctx is not None:
            self._grad_ctx.__exit__(*a)
        s.rec, s.train = self._old


def record(train_mode=True):
    return _Scope(True, train_mode)


def pause(train_mode=False):
    return _Scope(False, train_mode)


def train_mode():
    return _Scope(None, True)


def predict_mode():
    return _Scope(None, False)


def _register_leaf(arr):
    """Install the 'write' semantics hook: first gradient of a backward pass zeroes the buffer."""
    if arr._grad_hook is not None:
        arr._grad_hook.remove()
        arr._grad_hook = None
    if arr._grad_req not in ("write", "add"):
        return
    ref = weakref.ref(arr)
    t = arr._t
    write = arr._grad_req == "write"

    def hook(g):
        a = ref()
        if a is not None and a._bw_epoch != _bw_epoch[0]:
            a._bw_epoch = _bw_epoch[0]
            t._mx_fresh_grad = True          # a gradient arrived in this backward pass (NDArray._fresh_grad)
            if write and t.grad is not None:
                t.grad.zero_()
        return g

    arr._grad_hook = t.register_hook(hook)


def mark_variables(variables, gradients, grad_reqs="write"):
    if not isinstance(variables, (list, tuple)):
        variables, gradients = [variables], [gradients]
    if isinstance(grad_reqs, str):
        grad_reqs = [grad_reqs] * len(variables)
    from .ndarray.sparse import BaseSparseNDArray
    from .ndarray.ndarray import NDArray
    for v, g, r in zip(variables, gradients, grad_reqs):
        v.attach_grad(r)
        if isinstance(v, BaseSparseNDArray) or isinstance(g, BaseSparseNDArray):
            # sparse storage cannot alias the dense buffer autograd accumulates into: mirror it after every backward
            leaf = v.__dict__["_dense_leaf"]._t if isinstance(v, BaseSparseNDArray) else v._t

            def _mirror(p, g=g):
                if p.grad is not None:
                    g[:] = NDArray(p.grad.detach().clone(), g.context)
            if not hasattr(v, "_mirror_hooks"):
                try:
                    v._mirror_hooks = []
                except AttributeError:
                    pass
            h = leaf.register_post_accumulate_grad_hook(_mirror)
            getattr(v, "_mirror_hooks", []).append(h)
            continue
        v._t.grad = g._t


def backward(heads, head_grads=None, retain_graph=False, train_mode=True):
    if not isinstance(heads, (list, tuple)):
        heads = [heads]
    ts = [h._t for h in heads]
    if head_grads is None:
        gs = [torch.ones_like(t) for t in ts]
    else:
        if not isinstance(head_grads, (list, tuple)):
            head_grads = [head_grads]
        gs = [torch.ones_like(t) if g is None else g._t for t, g in zip(ts, head_grads)]
    # integer-valued heads (shape_array, argmax ...) are not differentiable: they contribute zero gradient
    keep = [i for i, t in enumerate(ts) if t.is_floating_point() or t.requires_grad]
    if len(keep) != len(ts):
        ts, gs = [ts[i] for i in keep], [gs[i] for i in keep]
        if not ts:
            return
    for t in ts:
        if getattr(t, "_mx_graph_freed", False):
            from .base import MXNetError
            raise MXNetError("the computation graph of this array has already been freed by a previous backward(); pass "
                             "retain_graph=True to the first call to differentiate it again")
    _bw_epoch[0] += 1
    torch.autograd.backward(ts, gs, retain_graph=retain_graph)
    if not retain_graph:
        for t in ts:
            if t.grad_fn is not None:
                t._mx_graph_freed = True


def grad(heads, variables, head_grads=None, retain_graph=None, create_graph=False, train_mode=True):
    from .ndarray.ndarray import NDArray
    single = not isinstance(variables, (list, tuple))
    if not isinstance(heads, (list, tuple)):
        heads = [heads]
    vs = [variables] if single else list(variables)
    ts = [h._t for h in heads]
    gs = None
    if head_grads is not None:
        if not isinstance(head_grads, (list, tuple)):
            head_grads = [head_grads]
        gs = [g._t for g in head_grads]
    else:
        gs = [torch.ones_like(t) for t in ts]
    rs = torch.autograd.grad(ts, [v._t for v in vs], gs, retain_graph=retain_graph if retain_graph is not None else create_graph,
                             create_graph=create_graph)
    out = [NDArray(r, v._ctx) for r, v in zip(rs, vs)]
    return out[0] if single else out


class Function(object):
    """User-defined differentiable function (reference: autograd.Function, python/mxnet/autograd.py:364).

    Subclasses implement ``forward(*inputs)`` and ``backward(*output_grads)`` on
    NDArrays; ``save_for_backward`` stores tensors on ``self``.
    """

    def save_for_backward(self, *args):
        self.saved_tensors = args

    def forward(self, *inputs):
        raise NotImplementedError

    def backward(self, *output_grads):
        raise NotImplementedError

    def __call__(self, *inputs):
        from .ndarray.ndarray import NDArray
        outer = self
        ctx0 = inputs[0]._ctx

        class _Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, *ts):
                with pause():
                    outs = outer.forward(*[NDArray(t.detach(), ctx0) for t in ts])
                outer._single = not isinstance(outs, (list, tuple))
                outs = [outs] if outer._single else list(outs)
                return tuple(o._t for o in outs)

            @staticmethod
            def backward(ctx, *gs):
                with pause():
                    r = outer.backward(*[NDArray(g, ctx0) for g in gs])
                r = [r] if not isinstance(r, (list, tuple)) else list(r)
                return tuple(x._t if x is not None else None for x in r)

        res = _Fn.apply(*[i._t for i in inputs])
        outs = [NDArray(r, ctx0) for r in res]
        return outs[0] if self._single else outs
