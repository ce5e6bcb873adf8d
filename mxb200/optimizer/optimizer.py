"""Optimizers and the Updater (reference API: python/mxnet/optimizer/optimizer.py).

Semantics kept from the reference: per-index update counting feeding the lr scheduler
(optimizer.py:373-410), lr/wd multipliers with ``wd_mult = 0`` for every parameter whose name does not end
in ``_weight`` / ``_gamma`` (optimizer.py:362-371), ``rescale_grad`` / ``clip_gradient`` applied to the raw
gradient, multi-precision (fp32 master copy for low-precision weights, optimizer.py:207-235), and the
sign conventions of each update rule.  The dense SGD/NAG/Adam/LAMB updates run as
single fused CUDA kernels on GPU tensors (mxb200/ops/optim.py); everything else is written with
elementwise NDArray math.  LAMB and the dynamic-mini-batch utilities are additions (absent upstream).

Provenance: adapted from Apache MXNet (incubating) 1.3.1 ``python/mxnet/optimizer.py`` (Apache License 2.0).  The public API, argument
names and control flow follow the original so that user scripts keep working; the tensor maths runs on torch /
the mxb200 kernels.  This file is API-compatibility glue, not original B200 work ("Provenance").
"""
import math
import pickle

import numpy as np
import torch

from ..base import MXNetError
from ..ndarray import NDArray, zeros, op as ndop, sparse as _sp
from ..ndarray.ndarray import zeros_like

__all__ = ["Optimizer", "register", "create", "get_updater", "Updater", "SGD", "NAG", "Signum", "FTML", "LBSGD",
           "DCASGD", "SGLD", "ccSGD", "Adam", "AdaGrad", "RMSProp", "AdaDelta", "Ftrl", "Adamax", "Nadam", "LAMB",
           "LARS", "Test"]

_LOW_PRECISION = (torch.float16, torch.bfloat16)


def _dense_storage_roundtrip(method):
    """Optimizers are written against dense weights / states; a row_sparse (or csr) weight or state is updated through a
    dense copy and written back in place (the reference keeps such arrays sparse with every row present).  Sparse
    GRADIENTS are passed through untouched: lazy-update semantics depend on them."""
    import functools

    @functools.wraps(method)
    def wrapped(self, index, weight, grad, state):
        leaves = []

        def densify(x):
            if isinstance(x, _sp.BaseSparseNDArray):
                d = x.todense()
                leaves.append((x, d))
                return d
            if isinstance(x, (tuple, list)):
                return type(x)(densify(v) for v in x)
            return x
        w, st = densify(weight), densify(state)
        if not leaves:
            return method(self, index, weight, grad, state)
        r = method(self, index, w, grad, st)
        for orig, dense in leaves:
            orig[:] = dense
        return r
    return wrapped


class Optimizer(object):
    opt_registry = {}

    def __init__(self, rescale_grad=1.0, param_idx2name=None, wd=0.0, clip_gradient=None, learning_rate=0.01,
                 lr_scheduler=None, sym=None, begin_num_update=0, multi_precision=False, param_dict=None):
        self.rescale_grad = rescale_grad
        self.lr = learning_rate
        self.lr_scheduler = lr_scheduler
        if lr_scheduler is not None:
            self.lr_scheduler.base_lr = learning_rate
        self.wd = wd
        self.lr_mult = {}
        self.wd_mult = {}
        self.begin_num_update = begin_num_update
        self.num_update = begin_num_update
        self._index_update_count = {}
        self.clip_gradient = clip_gradient
        self.multi_precision = multi_precision
        if param_idx2name is None:
            param_idx2name = {}
        assert isinstance(param_idx2name, dict), "param_idx2name should be a dict of param indexes to names."
        self.idx2name = param_idx2name.copy()
        self.sym_info = (sym.attr_dict(), sym.list_arguments()) if sym is not None else ()
        self.param_dict = param_dict if param_dict else {}
        self.set_lr_mult({})
        self.set_wd_mult({})

    # registry ---------------------------------------------------------------------------------
    @staticmethod
    def register(klass):
        assert isinstance(klass, type)
        Optimizer.opt_registry[klass.__name__.lower()] = klass
        for meth in ("update", "update_multi_precision"):
            if meth in klass.__dict__:
                setattr(klass, meth, _dense_storage_roundtrip(klass.__dict__[meth]))
        return klass

    @staticmethod
    def create_optimizer(name, **kwargs):
        if name.lower() in Optimizer.opt_registry:
            return Optimizer.opt_registry[name.lower()](**kwargs)
        raise ValueError("Cannot find optimizer %s" % name)

    # state -------------------------------------------------------------------------------------
    def create_state(self, index, weight):
        return None

    def create_state_multi_precision(self, index, weight):
        if self.multi_precision and weight._t.dtype in _LOW_PRECISION:
            master = NDArray(weight._t.detach().float().clone(), weight.context)
            return (master, self.create_state(index, master))
        if weight._t.dtype in _LOW_PRECISION and not self.multi_precision:
            import warnings
            warnings.warn("Accumulating with low-precision weights in optimizer can lead to poor accuracy or slow "
                          "convergence. Consider using multi_precision=True option of the optimizer")
        return self.create_state(index, weight)

    def update(self, index, weight, grad, state):
        raise NotImplementedError()

    def update_multi_precision(self, index, weight, grad, state):
        if self.multi_precision and weight._t.dtype in _LOW_PRECISION:
            master, inner = state
            sparse_g = getattr(grad, "stype", "default") != "default"
            # a row_sparse gradient stays row_sparse (the lazy update touches its rows only)
            g32 = grad.astype("float32") if sparse_g else NDArray(grad._t.float(), grad.context)
            self.update(index, master, g32, inner)
            if getattr(weight, "stype", "default") != "default":
                weight[:] = NDArray(master._t.to(weight._t.dtype), weight.context)     # storage-aware write-back
            else:
                with torch.no_grad():
                    weight._t.copy_(master._t)
        else:
            self.update(index, weight, grad, state)

    # hyper-parameter plumbing ------------------------------------------------------------------
    def set_learning_rate(self, lr):
        if self.lr_scheduler is not None:
            raise UserWarning("LRScheduler of the optimizer has already been defined. Note that set_learning_rate can "
                              "mutate the value of the learning rate of the optimizer only when the LRScheduler of "
                              "the optimizer is undefined.")
        self.lr = lr

    @property
    def learning_rate(self):
        if self.lr_scheduler is not None:
            return self.lr_scheduler(self.num_update)
        return self.lr

    def set_lr_scale(self, args_lrscale):
        raise DeprecationWarning

    def set_lr_mult(self, args_lr_mult):
        self.lr_mult = {}
        if self.sym_info:
            attr, arg_names = self.sym_info
            for name in arg_names:
                if name in attr and "__lr_mult__" in attr[name]:
                    self.lr_mult[name] = float(attr[name]["__lr_mult__"])
        self.lr_mult.update(args_lr_mult)

    def set_wd_mult(self, args_wd_mult):
        self.wd_mult = {}
        for n in self.idx2name.values():
            if not (n.endswith("_weight") or n.endswith("_gamma")):
                self.wd_mult[n] = 0.0
        if self.sym_info:
            attr, arg_names = self.sym_info
            for name in arg_names:
                if name in attr and "__wd_mult__" in attr[name]:
                    self.wd_mult[name] = float(attr[name]["__wd_mult__"])
        self.wd_mult.update(args_wd_mult)

    def _update_count(self, index):
        if index not in self._index_update_count:
            self._index_update_count[index] = self.begin_num_update
        self._index_update_count[index] += 1
        self.num_update = max(self._index_update_count[index], self.num_update)

    def _get_lr(self, index):
        lr = self.lr_scheduler(self.num_update) if self.lr_scheduler is not None else self.lr
        if index in self.param_dict:
            lr *= self.param_dict[index].lr_mult
        elif index in self.lr_mult:
            lr *= self.lr_mult[index]
        elif index in self.idx2name:
            lr *= self.lr_mult.get(self.idx2name[index], 1.0)
        return lr

    def _get_wd(self, index):
        wd = self.wd
        if index in self.param_dict:
            wd *= self.param_dict[index].wd_mult
        elif index in self.wd_mult:
            wd *= self.wd_mult[index]
        elif index in self.idx2name:
            wd *= self.wd_mult.get(self.idx2name[index], 1.0)
        return wd

    def __getstate__(self):
        d = self.__dict__.copy()
        d["param_dict"] = {}
        return d


register = Optimizer.register
create = Optimizer.create_optimizer


def _clip_kw(opt):
    return -1.0 if opt.clip_gradient is None else float(opt.clip_gradient)


def _prep(opt, grad, weight=None, wd=0.0):
    """rescale -> clip on an fp32 copy of the gradient tensor (returns torch tensor)."""
    g = grad._t.float() * opt.rescale_grad
    if opt.clip_gradient is not None:
        g = g.clamp(-opt.clip_gradient, opt.clip_gradient)
    if weight is not None and wd:
        g = g + wd * weight._t.float()
    return g


def _assign(dst, val):
    with torch.no_grad():
        dst._t.copy_(val)


@register
class SGD(Optimizer):
    """SGD with momentum; lazy row-sparse updates when ``grad`` is row_sparse (optimizer.py:450-553)."""

    def __init__(self, momentum=0.0, lazy_update=True, **kwargs):
        super(SGD, self).__init__(**kwargs)
        self.momentum = momentum
        self.lazy_update = lazy_update

    def create_state(self, index, weight):
        if self.momentum == 0.0:
            return None
        return zeros_like(weight)

    def create_state_multi_precision(self, index, weight):
        if self.multi_precision and weight._t.dtype in _LOW_PRECISION:
            master = NDArray(weight._t.detach().float().clone(), weight.context)
            return (self.create_state(index, master), master)
        return super(SGD, self).create_state_multi_precision(index, weight)

    def _update_impl(self, index, weight, grad, state, multi_precision=False):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        # row_sparse gradients / weights: the update operators themselves implement the lazy (rows present in the gradient
        # only) and standard semantics (ndarray/op.py::_sparse_aware_update)
        kw = dict(rescale_grad=self.rescale_grad, clip_gradient=_clip_kw(self), lazy_update=self.lazy_update)
        if not multi_precision:
            if state is not None:
                ndop.sgd_mom_update(weight, grad, state, lr=lr, wd=wd, momentum=self.momentum, out=weight, **kw)
            else:
                ndop.sgd_update(weight, grad, lr=lr, wd=wd, out=weight, **kw)
        else:
            if state[0] is not None:
                ndop.mp_sgd_mom_update(weight, grad, state[0], state[1], lr=lr, wd=wd, momentum=self.momentum,
                                       out=weight, **kw)
            else:
                ndop.mp_sgd_update(weight, grad, state[1], lr=lr, wd=wd, out=weight, **kw)

    def update(self, index, weight, grad, state):
        self._update_impl(index, weight, grad, state, multi_precision=False)

    def update_multi_precision(self, index, weight, grad, state):
        use_mp = self.multi_precision and weight._t.dtype in _LOW_PRECISION
        self._update_impl(index, weight, grad, state, multi_precision=use_mp)


@register
class ccSGD(SGD):
    """Deprecated alias of SGD."""


@register
class NAG(Optimizer):
    """Nesterov accelerated SGD: state = mu*state + g + wd*w ; w -= lr*(g + wd*w + mu*state)
    (optimizer.py:925-973)."""

    def __init__(self, momentum=0.0, **kwargs):
        super(NAG, self).__init__(**kwargs)
        self.momentum = momentum

    def create_state(self, index, weight):
        if self.momentum == 0.0:
            return None
        return zeros_like(weight)

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        w, g = weight._t, grad._t
        if w.is_cuda and w.is_contiguous() and g.is_contiguous() and (state is None or state._t.is_contiguous()):
            from ..ops import optim as KO
            with torch.no_grad():
                KO.sgd_update(w, g, state._t if state is not None else None, None, lr, wd, self.momentum,
                              self.rescale_grad, self.clip_gradient, nesterov=True)
            return
        with torch.no_grad():
            gg = _prep(self, grad)
            if state is not None:
                gg = gg + wd * w.float()
                m = self.momentum * state._t.float() + gg
                state._t.copy_(m)
                w.copy_(w.float() - lr * (gg + self.momentum * m))
            else:
                w.copy_(w.float() - lr * (gg + wd * w.float()))


@register
class Signum(Optimizer):
    """signSGD / Signum (optimizer.py:555-620)."""

    def __init__(self, learning_rate=0.01, momentum=0.9, wd_lh=0.0, **kwargs):
        super(Signum, self).__init__(learning_rate=learning_rate, **kwargs)
        self.momentum = momentum
        self.wd_lh = wd_lh

    def create_state(self, index, weight):
        return zeros_like(weight) if self.momentum != 0.0 else None

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        kw = dict(rescale_grad=self.rescale_grad, clip_gradient=_clip_kw(self))
        if state is not None:
            ndop.signum_update(weight, grad, state, lr=lr, wd=wd, momentum=self.momentum, wd_lh=self.wd_lh,
                               out=weight, **kw)
        else:
            ndop.signsgd_update(weight, grad, lr=lr, wd=wd, out=weight, **kw)


@register
class FTML(Optimizer):
    """Follow the Moving Leader (optimizer.py:622-676)."""

    def __init__(self, beta1=0.6, beta2=0.999, epsilon=1e-8, **kwargs):
        super(FTML, self).__init__(**kwargs)
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon

    def create_state(self, index, weight):
        return (zeros_like(weight), zeros_like(weight), zeros_like(weight))

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        t = self._index_update_count[index]
        d, v, z = state
        ndop.ftml_update(weight, grad, d, v, z, lr=lr, beta1=self.beta1, beta2=self.beta2, epsilon=self.epsilon, t=t,
                         wd=wd, rescale_grad=self.rescale_grad, clip_grad=_clip_kw(self), out=weight)


@register
class LBSGD(Optimizer):
    """Large-batch SGD with warm-up strategies incl. LARS layer-wise rate (optimizer.py:678-866)."""

    def __init__(self, momentum=0.0, multi_precision=False, warmup_strategy="linear", warmup_epochs=5, batch_scale=1,
                 updates_per_epoch=32, begin_epoch=0, num_epochs=60, **kwargs):
        super(LBSGD, self).__init__(multi_precision=multi_precision, **kwargs)
        self.momentum = momentum
        self.warmup_strategy, self.warmup_epochs = warmup_strategy, warmup_epochs
        self.batch_scale, self.updates_per_epoch = batch_scale, updates_per_epoch
        self.init_updates = begin_epoch * updates_per_epoch
        self.num_epochs = num_epochs
        self.lbmult = 1
        self.cumgrads = {}
        self.adaptive = False
        self.admult = 1

    def create_state(self, index, weight):
        return zeros_like(weight) if self.momentum != 0.0 else None

    def _get_lbmult(self, nup):
        nwup = self.warmup_epochs * self.updates_per_epoch
        strategy = self.warmup_strategy
        maxmult = float(self.batch_scale)
        if nup >= nwup:
            return maxmult
        if nwup <= 1:
            return 1.0
        if strategy == "linear":
            return 1.0 + (maxmult - 1) * nup / nwup
        if strategy == "power2":
            return 1.0 + (maxmult - 1) * (nup * nup) / (nwup * nwup)
        if strategy == "sqrt":
            return 1.0 + (maxmult - 1) * math.sqrt(float(nup) / nwup)
        return 1.0

    def _get_lars(self, weight, g, wd):
        w2 = float((weight._t.float() ** 2).sum())
        g2 = float((g ** 2).sum())
        lars = math.sqrt(w2 / (g2 + wd * w2 + 1e-18))
        return min(max(lars, 0.01), 100.0)

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        cum = self.cumgrads.setdefault(index, {"cum_grad": None, "num_cums": 0})
        g = grad._t.float()
        cum["cum_grad"] = g.clone() if cum["cum_grad"] is None or cum["num_cums"] == 0 else cum["cum_grad"] + g
        cum["num_cums"] += 1
        if cum["num_cums"] < self.batch_scale:
            return
        g = cum["cum_grad"] / self.batch_scale
        cum["num_cums"] = 0
        if self.warmup_strategy == "lars":
            lbmult = self._get_lars(weight, g * self.rescale_grad, wd)
        else:
            lbmult = self._get_lbmult(self.num_update + self.init_updates)
        lr = lr * lbmult
        gnd = NDArray(g.to(grad._t.dtype), grad.context)
        kw = dict(rescale_grad=self.rescale_grad, clip_gradient=_clip_kw(self))
        if state is not None:
            ndop.sgd_mom_update(weight, gnd, state, lr=lr, wd=wd, momentum=self.momentum, out=weight, **kw)
        else:
            ndop.sgd_update(weight, gnd, lr=lr, wd=wd, out=weight, **kw)


@register
class LARS(Optimizer):
    """Layer-wise Adaptive Rate Scaling with momentum (You et al. 2017): the stand-alone form of LBSGD's
    ``warmup_strategy='lars'``."""

    def __init__(self, momentum=0.9, eta=0.001, epsilon=1e-9, **kwargs):
        super(LARS, self).__init__(**kwargs)
        self.momentum, self.eta, self.epsilon = momentum, eta, epsilon

    def create_state(self, index, weight):
        return zeros_like(weight)

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        with torch.no_grad():
            g = _prep(self, grad)
            w = weight._t.float()
            wn, gn = w.norm(), g.norm()
            trust = torch.where((wn > 0) & (gn > 0), self.eta * wn / (gn + wd * wn + self.epsilon), torch.ones_like(wn))
            m = self.momentum * state._t.float() + lr * trust * (g + wd * w)
            state._t.copy_(m)
            weight._t.copy_(w - m)


@register
class DCASGD(Optimizer):
    """Delay-compensated async SGD (optimizer.py:868-922)."""

    def __init__(self, momentum=0.0, lamda=0.04, **kwargs):
        super(DCASGD, self).__init__(**kwargs)
        self.momentum, self.lamda = momentum, lamda
        self.weight_previous = {}

    def create_state(self, index, weight):
        return (zeros_like(weight) if self.momentum != 0.0 else None, weight.copy())

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        mom, prev = state
        with torch.no_grad():
            g = _prep(self, grad)
            w = weight._t.float()
            upd = -lr * (g + wd * w + self.lamda * g * g * (w - prev._t.float()))
            if mom is not None:
                m = self.momentum * mom._t.float() + upd
                mom._t.copy_(m)
                upd = m
            prev._t.copy_(w)
            weight._t.copy_(w + upd)


@register
class SGLD(Optimizer):
    """Stochastic Gradient Langevin Dynamics (optimizer.py:975-1013)."""

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        with torch.no_grad():
            g = _prep(self, grad)
            w = weight._t.float()
            noise = torch.randn_like(w) * math.sqrt(lr)
            weight._t.copy_(w - lr / 2 * (g + wd * w) + noise)


@register
class Adam(Optimizer):
    """Adam; bias correction folded into lr: lr*sqrt(1-b2^t)/(1-b1^t) (optimizer.py:1020-1100)."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, lazy_update=True, **kwargs):
        super(Adam, self).__init__(learning_rate=learning_rate, **kwargs)
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon
        self.lazy_update = lazy_update

    def create_state(self, index, weight):
        return (zeros_like(weight), zeros_like(weight))

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        t = self._index_update_count[index]
        lr *= math.sqrt(1.0 - self.beta2 ** t) / (1.0 - self.beta1 ** t)
        mean, var = state
        ndop.adam_update(weight, grad, mean, var, lr=lr, beta1=self.beta1, beta2=self.beta2, epsilon=self.epsilon, wd=wd,
                         rescale_grad=self.rescale_grad, clip_gradient=_clip_kw(self), lazy_update=self.lazy_update, out=weight)


@register
class LAMB(Optimizer):
    """Layer-wise Adaptive Moments for Batch training (You et al. 2019).  Not in the reference tree
    (SURVEY.md section 0); provided because the BERT-base config requires it.  Two-phase fused kernels on GPU."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-6, lower_bound=None, upper_bound=None,
                 bias_correction=True, **kwargs):
        super(LAMB, self).__init__(learning_rate=learning_rate, **kwargs)
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon
        self.lower_bound, self.upper_bound = lower_bound, upper_bound
        self.bias_correction = bias_correction

    def create_state(self, index, weight):
        return (NDArray(torch.zeros_like(weight._t, dtype=torch.float32), weight.context),
                NDArray(torch.zeros_like(weight._t, dtype=torch.float32), weight.context))

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        t = self._index_update_count[index]
        m, v = state
        bc1 = 1.0 - self.beta1 ** t if self.bias_correction else 1.0
        bc2 = 1.0 - self.beta2 ** t if self.bias_correction else 1.0
        with torch.no_grad():
            g = _prep(self, grad)
            w = weight._t.float()
            m._t.mul_(self.beta1).add_(g, alpha=1 - self.beta1)
            v._t.mul_(self.beta2).addcmul_(g, g, value=1 - self.beta2)
            r = (m._t / bc1) / ((v._t / bc2).sqrt() + self.epsilon) + wd * w
            wn, rn = w.norm(), r.norm()
            if self.lower_bound is not None:
                wn = wn.clamp(min=self.lower_bound)
            if self.upper_bound is not None:
                wn = wn.clamp(max=self.upper_bound)
            ratio = torch.where((wn > 0) & (rn > 0), wn / rn, torch.ones_like(wn))
            weight._t.copy_(w - lr * ratio * r)


@register
class AdaGrad(Optimizer):
    """AdaGrad (optimizer.py:1102-1160): history += g^2 ; w -= lr*(g/sqrt(history+eps) + wd*w)."""

    def __init__(self, eps=1e-7, **kwargs):
        super(AdaGrad, self).__init__(**kwargs)
        self.float_stable_eps = eps

    def create_state(self, index, weight):
        return zeros_like(weight)

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        if isinstance(grad, _sp.RowSparseNDArray):
            ndop.get("_sparse_adagrad_update")(weight, grad, state, lr=lr, epsilon=self.float_stable_eps, wd=wd,
                                               rescale_grad=self.rescale_grad, clip_gradient=_clip_kw(self))
            return
        with torch.no_grad():
            g = _prep(self, grad)
            state._t.add_(g * g)
            w = weight._t.float()
            weight._t.copy_(w - lr * (g / (state._t + self.float_stable_eps).sqrt() + wd * w))


@register
class RMSProp(Optimizer):
    """RMSProp, plain (Tieleman & Hinton) or centered (Graves) (optimizer.py:1162-1240)."""

    def __init__(self, learning_rate=0.001, gamma1=0.9, gamma2=0.9, epsilon=1e-8, centered=False, clip_weights=None,
                 **kwargs):
        super(RMSProp, self).__init__(learning_rate=learning_rate, **kwargs)
        self.gamma1, self.gamma2, self.centered = gamma1, gamma2, centered
        self.epsilon, self.clip_weights = epsilon, clip_weights

    def create_state(self, index, weight):
        if self.centered:
            return (zeros_like(weight), zeros_like(weight), zeros_like(weight))
        return (zeros_like(weight),)

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        kw = dict(rescale_grad=self.rescale_grad, clip_gradient=_clip_kw(self),
                  clip_weights=-1.0 if self.clip_weights is None else self.clip_weights)
        if not self.centered:
            ndop.rmsprop_update(weight, grad, state[0], lr=lr, gamma1=self.gamma1, epsilon=self.epsilon, wd=wd,
                                out=weight, **kw)
        else:
            n, g, delta = state
            ndop.rmspropalex_update(weight, grad, n, g, delta, lr=lr, gamma1=self.gamma1, gamma2=self.gamma2,
                                    epsilon=self.epsilon, wd=wd, out=weight, **kw)


@register
class AdaDelta(Optimizer):
    """AdaDelta (optimizer.py:1242-1295)."""

    def __init__(self, rho=0.90, epsilon=1e-5, **kwargs):
        super(AdaDelta, self).__init__(**kwargs)
        self.rho, self.epsilon = rho, epsilon

    def create_state(self, index, weight):
        return (zeros_like(weight), zeros_like(weight))

    def update(self, index, weight, grad, state):
        self._update_count(index)
        wd = self._get_wd(index)
        acc_g, acc_d = state
        with torch.no_grad():
            g = _prep(self, grad)
            acc_g._t.copy_(self.rho * acc_g._t + (1 - self.rho) * g * g)
            delta = torch.sqrt(acc_d._t + self.epsilon) / torch.sqrt(acc_g._t + self.epsilon) * g
            acc_d._t.copy_(self.rho * acc_d._t + (1 - self.rho) * delta * delta)
            w = weight._t.float()
            weight._t.copy_(w - (delta + wd * w))


@register
class Ftrl(Optimizer):
    """FTRL-proximal (optimizer.py:1297-1365)."""

    def __init__(self, lamda1=0.01, learning_rate=0.1, beta=1, **kwargs):
        super(Ftrl, self).__init__(learning_rate=learning_rate, **kwargs)
        self.lamda1, self.beta = lamda1, beta

    def create_state(self, index, weight):
        return (zeros_like(weight), zeros_like(weight))  # z, n

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        z, n = state
        ndop.ftrl_update(weight, grad, z, n, lr=lr, lamda1=self.lamda1, beta=self.beta, wd=wd,
                         rescale_grad=self.rescale_grad, clip_gradient=_clip_kw(self), out=weight)


@register
class Adamax(Optimizer):
    """AdaMax: infinity-norm variant of Adam (optimizer.py:1367-1420)."""

    def __init__(self, learning_rate=0.002, beta1=0.9, beta2=0.999, **kwargs):
        super(Adamax, self).__init__(learning_rate=learning_rate, **kwargs)
        self.beta1, self.beta2 = beta1, beta2

    def create_state(self, index, weight):
        return (zeros_like(weight), zeros_like(weight))

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        t = self._index_update_count[index]
        lr /= (1.0 - self.beta1 ** t)
        m, u = state
        with torch.no_grad():
            g = grad._t.float() * self.rescale_grad + wd * weight._t.float()
            if self.clip_gradient is not None:
                g = g.clamp(-self.clip_gradient, self.clip_gradient)
            m._t.copy_(self.beta1 * m._t + (1 - self.beta1) * g)
            u._t.copy_(torch.maximum(self.beta2 * u._t, g.abs()))
            weight._t.copy_(weight._t.float() - lr * m._t / u._t)


@register
class Nadam(Optimizer):
    """Nesterov Adam (optimizer.py:1422-1487)."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, schedule_decay=0.004, **kwargs):
        super(Nadam, self).__init__(learning_rate=learning_rate, **kwargs)
        self.beta1, self.beta2, self.epsilon, self.schedule_decay = beta1, beta2, epsilon, schedule_decay
        self.m_schedule = 1.0

    def create_state(self, index, weight):
        return (zeros_like(weight), zeros_like(weight))

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        t = self._index_update_count[index]
        m, v = state
        with torch.no_grad():
            g = grad._t.float() * self.rescale_grad + wd * weight._t.float()
            if self.clip_gradient is not None:
                g = g.clamp(-self.clip_gradient, self.clip_gradient)
            mom_t = self.beta1 * (1.0 - 0.5 * (0.96 ** (t * self.schedule_decay)))
            mom_t1 = self.beta1 * (1.0 - 0.5 * (0.96 ** ((t + 1) * self.schedule_decay)))
            self.m_schedule = self.m_schedule * mom_t
            m_sched_next = self.m_schedule * mom_t1
            m._t.copy_(self.beta1 * m._t + (1 - self.beta1) * g)
            v._t.copy_(self.beta2 * v._t + (1 - self.beta2) * g * g)
            g_prime = g / (1.0 - self.m_schedule)
            m_prime = m._t / (1.0 - m_sched_next)
            v_prime = v._t / (1.0 - self.beta2 ** t)
            m_bar = (1.0 - mom_t) * g_prime + mom_t1 * m_prime
            weight._t.copy_(weight._t.float() - lr * m_bar / (v_prime.sqrt() + self.epsilon))


@register
class Test(Optimizer):
    """The test optimizer used by the kvstore closed-form tests: w += rescale_grad * g (optimizer.py:1489-1502)."""

    def create_state(self, index, weight):
        return zeros(weight.shape, weight.context)

    def update(self, index, weight, grad, state):
        with torch.no_grad():
            weight._t.add_(grad._t.to(weight._t.dtype) * self.rescale_grad)
            state._t.copy_(weight._t)


class Updater(object):
    """Applies an optimizer to (index, grad, weight) triples, lazily creating per-index state
    (optimizer.py:1504-1571)."""

    def __init__(self, optimizer):
        self.optimizer = optimizer
        self.states = {}
        self.states_synced = {}

    def __call__(self, index, grad, weight):
        if index not in self.states:
            self.states[index] = self.optimizer.create_state_multi_precision(index, weight)
            self.states_synced[index] = True
        elif not self.states_synced[index]:
            self.states[index] = self.sync_state_context(self.states[index], weight.context)
            self.states_synced[index] = True
        self.optimizer.update_multi_precision(index, weight, grad, self.states[index])

    def sync_state_context(self, state, context):
        if isinstance(state, NDArray):
            return state.as_in_context(context)
        if isinstance(state, (tuple, list)):
            synced = [self.sync_state_context(i, context) for i in state]
            return tuple(synced) if isinstance(state, tuple) else synced
        return state

    def set_states(self, states):
        states = pickle.loads(states)
        if isinstance(states, tuple) and len(states) == 2:
            self.states, self.optimizer = states
        else:
            self.states = states
        self.states_synced = dict.fromkeys(self.states.keys(), False)

    def get_states(self, dump_optimizer=False):
        return pickle.dumps((self.states, self.optimizer) if dump_optimizer else self.states)


def get_updater(optimizer):
    return Updater(optimizer)
