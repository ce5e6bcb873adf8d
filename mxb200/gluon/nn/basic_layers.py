"""Basic neural network layers (reference: python/mxnet/gluon/nn/basic_layers.py).

Provenance: adapted from Apache MXNet (incubating) 1.3.1 ``python/mxnet/gluon/nn/basic_layers.py`` (Apache License 2.0).  The public API, argument
names and control flow follow the original so that user scripts keep working; the tensor maths runs on torch /
the mxb200 kernels.  This file is API-compatibility glue, not original B200 work ("Provenance").
"""
import warnings

import numpy as np

from ..block import Block, HybridBlock
from ..utils import _indent
from ... import autograd
from ...base import numeric_types

__all__ = ["Sequential", "HybridSequential", "Dense", "Dropout", "Embedding", "BatchNorm", "InstanceNorm",
           "LayerNorm", "Flatten", "Lambda", "HybridLambda", "Activation"]


class Sequential(Block):
    def __init__(self, prefix=None, params=None):
        super(Sequential, self).__init__(prefix=prefix, params=params)

    def add(self, *blocks):
        for block in blocks:
            self.register_child(block)

    def forward(self, x):
        for block in self._children.values():
            x = block(x)
        return x

    def __repr__(self):
        s = "{name}(\n{modstr}\n)"
        modstr = "\n".join(["  ({key}): {block}".format(key=key, block=_indent(block.__repr__(), 2))
                            for key, block in self._children.items()])
        return s.format(name=self.__class__.__name__, modstr=modstr)

    def __getitem__(self, key):
        layers = list(self._children.values())[key]
        if isinstance(layers, list):
            net = type(self)(prefix=self._prefix)
            with net.name_scope():
                net.add(*layers)
            return net
        return layers

    def __len__(self):
        return len(self._children)

    def hybridize(self, active=True, **kwargs):
        if self._children and all(isinstance(c, HybridBlock) for c in self._children.values()):
            warnings.warn("All children of this Sequential layer '%s' are HybridBlocks. Consider using "
                          "HybridSequential for the best performance." % self.prefix, stacklevel=2)
        super(Sequential, self).hybridize(active, **kwargs)


class HybridSequential(HybridBlock):
    def __init__(self, prefix=None, params=None):
        super(HybridSequential, self).__init__(prefix=prefix, params=params)

    def add(self, *blocks):
        for block in blocks:
            self.register_child(block)

    def hybrid_forward(self, F, x):
        for block in self._children.values():
            x = block(x)
        return x

    def __repr__(self):
        s = "{name}(\n{modstr}\n)"
        modstr = "\n".join(["  ({key}): {block}".format(key=key, block=_indent(block.__repr__(), 2))
                            for key, block in self._children.items()])
        return s.format(name=self.__class__.__name__, modstr=modstr)

    def __getitem__(self, key):
        layers = list(self._children.values())[key]
        if isinstance(layers, list):
            net = type(self)(prefix=self._prefix)
            with net.name_scope():
                net.add(*layers)
            return net
        return layers

    def __len__(self):
        return len(self._children)


class Dense(HybridBlock):
    """output = activation(dot(input, weight.T) + bias); weight (units, in_units)."""

    def __init__(self, units, activation=None, use_bias=True, flatten=True, dtype="float32", weight_initializer=None,
                 bias_initializer="zeros", in_units=0, **kwargs):
        super(Dense, self).__init__(**kwargs)
        self._flatten = flatten
        with self.name_scope():
            self._units = units
            self._in_units = in_units
            self.weight = self.params.get("weight", shape=(units, in_units), init=weight_initializer, dtype=dtype,
                                          allow_deferred_init=True)
            if use_bias:
                self.bias = self.params.get("bias", shape=(units,), init=bias_initializer, dtype=dtype,
                                            allow_deferred_init=True)
            else:
                self.bias = None
            if activation is not None:
                self.act = Activation(activation, prefix=activation + "_")
            else:
                self.act = None

    def _infer_param_shapes(self, x, *args):
        in_units = int(np.prod(x.shape[1:])) if self._flatten else x.shape[-1]
        self.weight.shape = (self._units, in_units)

    def hybrid_forward(self, F, x, weight, bias=None):
        act = F.FullyConnected(x, weight, bias, no_bias=bias is None, num_hidden=self._units, flatten=self._flatten,
                               name="fwd")
        if self.act is not None:
            act = self.act(act)
        return act

    def __repr__(self):
        s = "{name}({layout}, {act})"
        shape = self.weight.shape
        return s.format(name=self.__class__.__name__, act=self.act if self.act else "linear",
                        layout="{0} -> {1}".format(shape[1] if shape[1] else None, shape[0]))


class Activation(HybridBlock):
    def __init__(self, activation, **kwargs):
        self._act_type = activation
        super(Activation, self).__init__(**kwargs)

    def _alias(self):
        return self._act_type

    def hybrid_forward(self, F, x):
        return F.Activation(x, act_type=self._act_type, name="fwd")

    def __repr__(self):
        return "{name}({_act_type})".format(name=self.__class__.__name__, **self.__dict__)


class Dropout(HybridBlock):
    def __init__(self, rate, axes=(), **kwargs):
        super(Dropout, self).__init__(**kwargs)
        self._rate = rate
        self._axes = axes

    def hybrid_forward(self, F, x):
        return F.Dropout(x, p=self._rate, axes=self._axes, name="fwd")

    def __repr__(self):
        return "{name}(p = {_rate}, axes={_axes})".format(name=self.__class__.__name__, **self.__dict__)


class BatchNorm(HybridBlock):
    """Batch normalization (nn/batch_norm-inl.h semantics: momentum multiplies the *running* value)."""

    def __init__(self, axis=1, momentum=0.9, epsilon=1e-5, center=True, scale=True, use_global_stats=False,
                 beta_initializer="zeros", gamma_initializer="ones", running_mean_initializer="zeros",
                 running_variance_initializer="ones", in_channels=0, **kwargs):
        super(BatchNorm, self).__init__(**kwargs)
        self._kwargs = {"axis": axis, "eps": epsilon, "momentum": momentum, "fix_gamma": not scale,
                        "use_global_stats": use_global_stats}
        self._axis = axis
        if in_channels != 0:
            self.in_channels = in_channels
        self.gamma = self.params.get("gamma", grad_req="write" if scale else "null", shape=(in_channels,),
                                     init=gamma_initializer, allow_deferred_init=True, differentiable=scale)
        self.beta = self.params.get("beta", grad_req="write" if center else "null", shape=(in_channels,),
                                    init=beta_initializer, allow_deferred_init=True, differentiable=center)
        self.running_mean = self.params.get("running_mean", grad_req="null", shape=(in_channels,),
                                            init=running_mean_initializer, allow_deferred_init=True,
                                            differentiable=False)
        self.running_var = self.params.get("running_var", grad_req="null", shape=(in_channels,),
                                           init=running_variance_initializer, allow_deferred_init=True,
                                           differentiable=False)

    def _infer_param_shapes(self, x, *args):
        c = x.shape[self._axis]
        for p in (self.gamma, self.beta, self.running_mean, self.running_var):
            p.shape = (c,)

    def cast(self, dtype):
        from ...base import dtype_name
        if dtype_name(dtype) in ("float16", "bfloat16"):
            dtype = "float32"   # BN params/stats stay fp32 for low-precision nets (block.py BatchNorm.cast)
        super(BatchNorm, self).cast(dtype)

    def hybrid_forward(self, F, x, gamma, beta, running_mean, running_var):
        return F.BatchNorm(x, gamma, beta, running_mean, running_var, name="fwd", **self._kwargs)

    def __repr__(self):
        s = "{name}({content}"
        in_channels = self.gamma.shape[0]
        s += ", in_channels={0}".format(in_channels if in_channels else None)
        s += ")"
        return s.format(name=self.__class__.__name__,
                        content=", ".join(["=".join([k, v.__repr__()]) for k, v in self._kwargs.items()]))


class Embedding(HybridBlock):
    def __init__(self, input_dim, output_dim, dtype="float32", weight_initializer=None, sparse_grad=False, **kwargs):
        super(Embedding, self).__init__(**kwargs)
        grad_stype = "row_sparse" if sparse_grad else "default"
        self._kwargs = {"input_dim": input_dim, "output_dim": output_dim, "dtype": dtype, "sparse_grad": sparse_grad}
        self.weight = self.params.get("weight", shape=(input_dim, output_dim), init=weight_initializer, dtype=dtype,
                                      allow_deferred_init=True, grad_stype=grad_stype)

    def hybrid_forward(self, F, x, weight):
        return F.Embedding(x, weight, name="fwd", **self._kwargs)

    def __repr__(self):
        return "{block_name}({input_dim} -> {output_dim}, {dtype})".format(block_name=self.__class__.__name__, **self._kwargs)


class Flatten(HybridBlock):
    def hybrid_forward(self, F, x):
        return F.Flatten(x)

    def __repr__(self):
        return self.__class__.__name__


class InstanceNorm(HybridBlock):
    def __init__(self, axis=1, epsilon=1e-5, center=True, scale=False, beta_initializer="zeros",
                 gamma_initializer="ones", in_channels=0, **kwargs):
        super(InstanceNorm, self).__init__(**kwargs)
        self._kwargs = {"eps": epsilon, "axis": axis, "center": center, "scale": scale}
        self._axis = axis
        self._epsilon = epsilon
        self.gamma = self.params.get("gamma", grad_req="write" if scale else "null", shape=(in_channels,),
                                     init=gamma_initializer, allow_deferred_init=True)
        self.beta = self.params.get("beta", grad_req="write" if center else "null", shape=(in_channels,),
                                    init=beta_initializer, allow_deferred_init=True)

    def _infer_param_shapes(self, x, *args):
        self.gamma.shape = (x.shape[self._axis],)
        self.beta.shape = (x.shape[self._axis],)

    def hybrid_forward(self, F, x, gamma, beta):
        if self._axis == 1:
            return F.InstanceNorm(x, gamma, beta, name="fwd", eps=self._epsilon)
        x = x.swapaxes(1, self._axis)
        return F.InstanceNorm(x, gamma, beta, name="fwd", eps=self._epsilon).swapaxes(1, self._axis)


class LayerNorm(HybridBlock):
    def __init__(self, axis=-1, epsilon=1e-5, center=True, scale=True, beta_initializer="zeros",
                 gamma_initializer="ones", in_channels=0, prefix=None, params=None):
        super(LayerNorm, self).__init__(prefix=prefix, params=params)
        self._kwargs = {"eps": epsilon, "axis": axis, "center": center, "scale": scale}
        self._axis = axis
        self._epsilon = epsilon
        self._center = center
        self._scale = scale
        self.gamma = self.params.get("gamma", grad_req="write" if scale else "null", shape=(in_channels,),
                                     init=gamma_initializer, allow_deferred_init=True)
        self.beta = self.params.get("beta", grad_req="write" if center else "null", shape=(in_channels,),
                                    init=beta_initializer, allow_deferred_init=True)

    def _infer_param_shapes(self, x, *args):
        self.gamma.shape = (x.shape[self._axis],)
        self.beta.shape = (x.shape[self._axis],)

    def hybrid_forward(self, F, data, gamma, beta):
        return F.LayerNorm(data, gamma=gamma, beta=beta, axis=self._axis, eps=self._epsilon)

    def __repr__(self):
        s = "{name}({content}"
        in_channels = self.gamma.shape[0]
        s += ", in_channels={0}".format(in_channels)
        s += ")"
        return s.format(name=self.__class__.__name__,
                        content=", ".join(["=".join([k, v.__repr__()]) for k, v in self._kwargs.items()]))


class Lambda(Block):
    def __init__(self, function, prefix=None):
        super(Lambda, self).__init__(prefix=prefix)
        if isinstance(function, str):
            from ... import ndarray as nd
            assert hasattr(nd, function), "Function name %s is not found in ndarray." % function
            self._func_impl = getattr(nd, function)
        elif callable(function):
            self._func_impl = function
        else:
            raise ValueError("Unrecognized function in lambda: {} of type {}".format(function, type(function)))

    def forward(self, *args):
        return self._func_impl(*args)

    def __repr__(self):
        return "{name}({function})".format(name=self.__class__.__name__, function=self._func_impl.__name__)


class HybridLambda(HybridBlock):
    def __init__(self, function, prefix=None):
        super(HybridLambda, self).__init__(prefix=prefix)
        if isinstance(function, str):
            from ... import ndarray as nd
            assert hasattr(nd, function), "Function name %s is not found in symbol/ndarray." % function
            self._func = lambda F, *args: getattr(F, function)(*args)
            self._func_name = function
        elif callable(function):
            self._func = function
            self._func_name = function.__name__
        else:
            raise ValueError("Unrecognized function in lambda: {} of type {}".format(function, type(function)))

    def hybrid_forward(self, F, x, *args):
        return self._func(F, x, *args)

    def __repr__(self):
        return "{name}({function})".format(name=self.__class__.__name__, function=self._func_name)
