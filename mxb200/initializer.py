"""Weight initializers (reference: python/mxnet/initializer.py).

``init(InitDesc(name, attrs), arr)`` dispatches on the parameter-name suffix exactly like the reference
(_weight / _bias / _gamma / _beta / moving_* ... initializer.py:120-170) so Module and Gluon get the same
defaults: bias/beta/moving_mean -> 0, gamma/moving_var -> 1, weights -> the chosen scheme.

Provenance: adapted from Apache MXNet (incubating) 1.3.1 ``python/mxnet/initializer.py`` (Apache License 2.0).  The public API, argument
names and control flow follow the original so that user scripts keep working; the tensor maths runs on torch /
the mxb200 kernels.  This file is API-compatibility glue, not original B200 work ("Provenance").
"""
import json
import math
import re

import numpy as np
import torch

from .ndarray import NDArray

__all__ = ["InitDesc", "Initializer", "register", "create", "Zero", "One", "Constant", "Uniform", "Normal",
           "Orthogonal", "Xavier", "MSRAPrelu", "Bilinear", "LSTMBias", "FusedRNN", "Load", "Mixed"]

_REG = {}


class InitDesc(str):
    def __new__(cls, name, attrs=None, global_init=None):
        ret = super(InitDesc, cls).__new__(cls, name)
        ret.attrs = attrs or {}
        ret.global_init = global_init
        return ret


def register(klass):
    _REG[klass.__name__.lower()] = klass
    return klass


def create(name, **kwargs):
    if isinstance(name, Initializer):
        return name
    if name is None:
        return None
    if isinstance(name, str) and name.startswith("["):
        k, kw = json.loads(name)
        return _REG[k.lower()](**kw)
    return _REG[name.lower()](**kwargs)


class Initializer(object):
    def __init__(self, **kwargs):
        self._kwargs = kwargs
        self._verbose = False
        self._print_func = None

    def set_verbosity(self, verbose=False, print_func=None):
        self._verbose = verbose
        self._print_func = print_func or (lambda x: str(float(x.norm().asscalar()) / math.sqrt(x.size)))
        return self

    def dumps(self):
        return json.dumps([self.__class__.__name__.lower(), self._kwargs])

    def __call__(self, desc, arr):
        if not isinstance(desc, InitDesc):
            desc = InitDesc(str(desc))
        if desc.global_init is None:
            desc.global_init = self
        init = desc.attrs.get("__init__", "")
        if init:
            create(init)._init_weight(desc, arr)
            return
        name = str(desc)
        if name.endswith("weight"):
            self._init_weight(desc, arr)
        elif name.endswith("bias"):
            self._init_bias(desc, arr)
        elif name.endswith("gamma"):
            self._init_gamma(desc, arr)
        elif name.endswith("beta"):
            self._init_beta(desc, arr)
        elif name.endswith("moving_mean") or name.endswith("running_mean"):
            self._init_zero(desc, arr)
        elif name.endswith("moving_var") or name.endswith("running_var"):
            self._init_one(desc, arr)
        elif name.endswith("moving_inv_var") or name.endswith("moving_avg"):
            self._init_zero(desc, arr)
        elif name.endswith("min") or name.endswith("max"):
            self._init_zero(desc, arr)
        else:
            self._init_default(desc, arr)

    def _fill(self, arr, t):
        with torch.no_grad():
            arr._t.copy_(t.to(arr._t.dtype))

    def _init_zero(self, _, arr):
        arr[:] = 0.0

    def _init_one(self, _, arr):
        arr[:] = 1.0

    def _init_bias(self, _, arr):
        arr[:] = 0.0

    def _init_gamma(self, _, arr):
        arr[:] = 1.0

    def _init_beta(self, _, arr):
        arr[:] = 0.0

    def _init_weight(self, name, arr):
        raise NotImplementedError("Must override it")

    def _init_default(self, name, _):
        raise ValueError("Unknown initialization pattern for %s. Default initialization is now limited to "
                         '"weight", "bias", "gamma" (1.0), and "beta" (0.0). Please use mx.sym.Variable(init=...) '
                         "to set initialization pattern" % name)

    def __eq__(self, other):
        return isinstance(other, Initializer) and self.__class__ is other.__class__ and self._kwargs == other._kwargs

    __hash__ = object.__hash__


@register
class Load(object):
    """Initialise from a dict / .params file, falling back to ``default_init``."""

    def __init__(self, param, default_init=None, verbose=False):
        if isinstance(param, str):
            from .ndarray import load
            param = load(param)
        self.param = {}
        for name, arr in param.items():
            self.param[name[4:] if name.startswith("arg:") or name.startswith("aux:") else name] = arr
        self.default_init = default_init
        self.verbose = verbose

    def __call__(self, name, arr):
        if name in self.param:
            assert arr.shape == self.param[name].shape, "Parameter %s cannot be initialized from loading. Shape mismatch" % name
            arr[:] = self.param[name]
        else:
            assert self.default_init is not None, "Cannot Initialize %s. Not found in loaded param and no default" % name
            self.default_init(name, arr)


@register
class Mixed(object):
    def __init__(self, patterns, initializers):
        assert len(patterns) == len(initializers)
        self.map = list(zip([re.compile(p) for p in patterns], initializers))

    def __call__(self, name, arr):
        for prog, init in self.map:
            if prog.match(name):
                init(name, arr)
                return
        raise ValueError("Parameter name %s did not match any pattern. Consider adding a \".*\" pattern" % name)


@register
class Zero(Initializer):
    def _init_weight(self, _, arr):
        arr[:] = 0.0


@register
class One(Initializer):
    def _init_weight(self, _, arr):
        arr[:] = 1.0


_REG["zeros"] = Zero
_REG["ones"] = One


@register
class Constant(Initializer):
    def __init__(self, value):
        super(Constant, self).__init__(value=value)
        self.value = value

    def _init_weight(self, _, arr):
        arr[:] = self.value

    _init_default = _init_weight
    _init_bias = _init_weight
    _init_gamma = _init_weight
    _init_beta = _init_weight


@register
class Uniform(Initializer):
    def __init__(self, scale=0.07):
        super(Uniform, self).__init__(scale=scale)
        self.scale = scale

    def _init_weight(self, _, arr):
        self._fill(arr, torch.empty(arr.shape).uniform_(-self.scale, self.scale))


@register
class Normal(Initializer):
    def __init__(self, sigma=0.01):
        super(Normal, self).__init__(sigma=sigma)
        self.sigma = sigma

    def _init_weight(self, _, arr):
        self._fill(arr, torch.empty(arr.shape).normal_(0, self.sigma))


@register
class Orthogonal(Initializer):
    def __init__(self, scale=1.414, rand_type="uniform"):
        super(Orthogonal, self).__init__(scale=scale, rand_type=rand_type)
        self.scale, self.rand_type = scale, rand_type

    def _init_weight(self, _, arr):
        nout = arr.shape[0]
        nin = int(np.prod(arr.shape[1:]))
        tmp = np.random.uniform(-1.0, 1.0, (nout, nin)) if self.rand_type == "uniform" else np.random.normal(0.0, 1.0, (nout, nin))
        u, _s, v = np.linalg.svd(tmp, full_matrices=False)
        q = u if u.shape == tmp.shape else v
        self._fill(arr, torch.from_numpy(self.scale * q.reshape(arr.shape)).float())


@register
class Xavier(Initializer):
    def __init__(self, rnd_type="uniform", factor_type="avg", magnitude=3):
        super(Xavier, self).__init__(rnd_type=rnd_type, factor_type=factor_type, magnitude=magnitude)
        self.rnd_type, self.factor_type, self.magnitude = rnd_type, factor_type, float(magnitude)

    def _init_weight(self, name, arr):
        shape = arr.shape
        hw_scale = 1.0
        if len(shape) < 2:
            raise ValueError("Xavier initializer cannot be applied to vector {0}. It requires at least 2D.".format(name))
        if len(shape) > 2:
            hw_scale = np.prod(shape[2:])
        fan_in, fan_out = shape[1] * hw_scale, shape[0] * hw_scale
        factor = {"avg": (fan_in + fan_out) / 2.0, "in": fan_in, "out": fan_out}[self.factor_type]
        scale = np.sqrt(self.magnitude / factor)
        if self.rnd_type == "uniform":
            self._fill(arr, torch.empty(shape).uniform_(-scale, scale))
        elif self.rnd_type == "gaussian":
            self._fill(arr, torch.empty(shape).normal_(0, scale))
        else:
            raise ValueError("Unknown random type")


@register
class MSRAPrelu(Xavier):
    def __init__(self, factor_type="avg", slope=0.25):
        magnitude = 2.0 / (1 + slope ** 2)
        super(MSRAPrelu, self).__init__("gaussian", factor_type, magnitude)
        self._kwargs = {"factor_type": factor_type, "slope": slope}


@register
class Bilinear(Initializer):
    def _init_weight(self, _, arr):
        weight = np.zeros(int(np.prod(arr.shape)), dtype="float32")
        shape = arr.shape
        f = np.ceil(shape[3] / 2.0)
        c = (2 * f - 1 - f % 2) / (2.0 * f)
        for i in range(int(np.prod(shape))):
            x = i % shape[3]
            y = (i // shape[3]) % shape[2]
            weight[i] = (1 - abs(x / f - c)) * (1 - abs(y / f - c))
        self._fill(arr, torch.from_numpy(weight.reshape(shape)))


@register
class LSTMBias(Initializer):
    def __init__(self, forget_bias=1.0):
        super(LSTMBias, self).__init__(forget_bias=forget_bias)
        self.forget_bias = forget_bias

    def _init_weight(self, name, arr):
        arr[:] = 0.0
        n = arr.shape[0] // 4
        with torch.no_grad():
            arr._t[n:2 * n] = self.forget_bias

    _init_bias = _init_weight


@register
class FusedRNN(Initializer):
    def __init__(self, init, num_hidden, num_layers, mode, bidirectional=False, forget_bias=1.0):
        if isinstance(init, str):
            init = create(init)
        super(FusedRNN, self).__init__(init=init.dumps() if init is not None else None, num_hidden=num_hidden,
                                       num_layers=num_layers, mode=mode, bidirectional=bidirectional,
                                       forget_bias=forget_bias)
        self._init = init

        self._num_hidden, self._num_layers, self._mode = num_hidden, num_layers, mode
        self._bidirectional, self._forget_bias = bidirectional, forget_bias

    def _init_weight(self, desc, arr):
        """Initialise the packed parameter vector gate by gate: weights with ``init`` (or the global initializer), biases
        with their own rule and the LSTM forget-gate bias with ``forget_bias`` (initializer.py:FusedRNN of the reference)."""
        from .rnn import rnn_cell
        cell = rnn_cell.FusedRNNCell(self._num_hidden, self._num_layers, self._mode, self._bidirectional,
                                     forget_bias=self._forget_bias, prefix="")
        args = cell.unpack_weights({"parameters": arr})
        for name in args:
            arg_desc = InitDesc(name, global_init=getattr(desc, "global_init", None))
            if self._mode == "lstm" and name.endswith("_f_bias"):
                args[name][:] = self._forget_bias
            elif self._init is None:
                g = getattr(desc, "global_init", None)
                (g if g is not None else Uniform(0.07))(arg_desc, args[name])
            else:
                self._init(arg_desc, args[name])
        arr[:] = cell.pack_weights(args)["parameters"]
