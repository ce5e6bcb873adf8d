"""Convolution / pooling layers (reference: python/mxnet/gluon/nn/conv_layers.py)."""
import numpy as np

from ..block import HybridBlock
from .basic_layers import Activation

__all__ = ["Conv1D", "Conv2D", "Conv3D", "Conv1DTranspose", "Conv2DTranspose", "Conv3DTranspose", "MaxPool1D",
           "MaxPool2D", "MaxPool3D", "AvgPool1D", "AvgPool2D", "AvgPool3D", "GlobalMaxPool1D", "GlobalMaxPool2D",
           "GlobalMaxPool3D", "GlobalAvgPool1D", "GlobalAvgPool2D", "GlobalAvgPool3D", "ReflectionPad2D"]


def _tup(v, n):
    return (v,) * n if isinstance(v, (int, np.integer)) else tuple(v)


class _Conv(HybridBlock):
    def __init__(self, channels, kernel_size, strides, padding, dilation, groups, layout, in_channels=0,
                 activation=None, use_bias=True, weight_initializer=None, bias_initializer="zeros",
                 op_name="Convolution", adj=None, prefix=None, params=None):
        super(_Conv, self).__init__(prefix=prefix, params=params)
        with self.name_scope():
            self._channels = channels
            self._in_channels = in_channels
            nd_ = len(kernel_size)
            self._kwargs = {"kernel": kernel_size, "stride": _tup(strides, nd_), "dilate": _tup(dilation, nd_),
                            "pad": _tup(padding, nd_), "num_filter": channels, "num_group": groups,
                            "no_bias": not use_bias, "layout": layout}
            if adj is not None:
                self._kwargs["adj"] = adj
            self._op_name = op_name
            self._layout = layout
            self._groups = groups
            self._transpose = op_name == "Deconvolution"
            wshape = self._wshape(in_channels)
            self.weight = self.params.get("weight", shape=wshape, init=weight_initializer, allow_deferred_init=True)
            if use_bias:
                self.bias = self.params.get("bias", shape=(channels,), init=bias_initializer, allow_deferred_init=True)
            else:
                self.bias = None
            self.act = Activation(activation, prefix=activation + "_") if activation is not None else None

    def _wshape(self, in_channels):
        k = tuple(self._kwargs["kernel"])
        if self._transpose:
            return (in_channels, self._channels // self._groups) + k
        return (self._channels, in_channels // self._groups if in_channels else 0) + k

    def _channel_axis(self):
        return self._layout.find("C")

    def _infer_param_shapes(self, x, *args):
        self.weight.shape = self._wshape(x.shape[self._channel_axis()])

    def hybrid_forward(self, F, x, weight, bias=None):
        kw = dict(self._kwargs)
        layout = kw.pop("layout")
        channels_last = layout is not None and layout.endswith("C")
        if channels_last:
            nd_ = len(kw["kernel"])
            perm_in = (0, nd_ + 1) + tuple(range(1, nd_ + 1))
            perm_out = (0,) + tuple(range(2, nd_ + 2)) + (1,)
            x = F.transpose(x, axes=perm_in)
        act = getattr(F, self._op_name)(x, weight, bias, name="fwd", **kw)
        if channels_last:
            act = F.transpose(act, axes=perm_out)
        if self.act is not None:
            act = self.act(act)
        return act

    def _alias(self):
        return "conv"

    def __repr__(self):
        s = "{name}({mapping}, kernel_size={kernel}, stride={stride}"
        shape = self.weight.shape
        return (s + ")").format(name=self.__class__.__name__,
                                mapping="{0} -> {1}".format(shape[1] if shape[1] else None, shape[0]), **self._kwargs)


class Conv1D(_Conv):
    def __init__(self, channels, kernel_size, strides=1, padding=0, dilation=1, groups=1, layout="NCW", activation=None,
                 use_bias=True, weight_initializer=None, bias_initializer="zeros", in_channels=0, **kwargs):
        super(Conv1D, self).__init__(channels, _tup(kernel_size, 1), strides, padding, dilation, groups, layout,
                                     in_channels, activation, use_bias, weight_initializer, bias_initializer, **kwargs)


class Conv2D(_Conv):
    def __init__(self, channels, kernel_size, strides=(1, 1), padding=(0, 0), dilation=(1, 1), groups=1, layout="NCHW",
                 activation=None, use_bias=True, weight_initializer=None, bias_initializer="zeros", in_channels=0,
                 **kwargs):
        super(Conv2D, self).__init__(channels, _tup(kernel_size, 2), strides, padding, dilation, groups, layout,
                                     in_channels, activation, use_bias, weight_initializer, bias_initializer, **kwargs)


class Conv3D(_Conv):
    def __init__(self, channels, kernel_size, strides=(1, 1, 1), padding=(0, 0, 0), dilation=(1, 1, 1), groups=1,
                 layout="NCDHW", activation=None, use_bias=True, weight_initializer=None, bias_initializer="zeros",
                 in_channels=0, **kwargs):
        super(Conv3D, self).__init__(channels, _tup(kernel_size, 3), strides, padding, dilation, groups, layout,
                                     in_channels, activation, use_bias, weight_initializer, bias_initializer, **kwargs)


class Conv1DTranspose(_Conv):
    def __init__(self, channels, kernel_size, strides=1, padding=0, output_padding=0, dilation=1, groups=1,
                 layout="NCW", activation=None, use_bias=True, weight_initializer=None, bias_initializer="zeros",
                 in_channels=0, **kwargs):
        super(Conv1DTranspose, self).__init__(channels, _tup(kernel_size, 1), strides, padding, dilation, groups,
                                              layout, in_channels, activation, use_bias, weight_initializer,
                                              bias_initializer, op_name="Deconvolution", adj=_tup(output_padding, 1),
                                              **kwargs)


class Conv2DTranspose(_Conv):
    def __init__(self, channels, kernel_size, strides=(1, 1), padding=(0, 0), output_padding=(0, 0), dilation=(1, 1),
                 groups=1, layout="NCHW", activation=None, use_bias=True, weight_initializer=None,
                 bias_initializer="zeros", in_channels=0, **kwargs):
        super(Conv2DTranspose, self).__init__(channels, _tup(kernel_size, 2), strides, padding, dilation, groups,
                                              layout, in_channels, activation, use_bias, weight_initializer,
                                              bias_initializer, op_name="Deconvolution", adj=_tup(output_padding, 2),
                                              **kwargs)


class Conv3DTranspose(_Conv):
    def __init__(self, channels, kernel_size, strides=(1, 1, 1), padding=(0, 0, 0), output_padding=(0, 0, 0),
                 dilation=(1, 1, 1), groups=1, layout="NCDHW", activation=None, use_bias=True, weight_initializer=None,
                 bias_initializer="zeros", in_channels=0, **kwargs):
        super(Conv3DTranspose, self).__init__(channels, _tup(kernel_size, 3), strides, padding, dilation, groups,
                                              layout, in_channels, activation, use_bias, weight_initializer,
                                              bias_initializer, op_name="Deconvolution", adj=_tup(output_padding, 3),
                                              **kwargs)


class _Pooling(HybridBlock):
    def __init__(self, pool_size, strides, padding, ceil_mode, global_pool, pool_type, layout=None,
                 count_include_pad=None, **kwargs):
        super(_Pooling, self).__init__(**kwargs)
        if strides is None:
            strides = pool_size
        nd_ = len(pool_size)
        self._layout = layout
        self._kwargs = {"kernel": pool_size, "stride": _tup(strides, nd_), "pad": _tup(padding, nd_),
                        "global_pool": global_pool, "pool_type": pool_type,
                        "pooling_convention": "full" if ceil_mode else "valid"}
        if count_include_pad is not None:
            self._kwargs["count_include_pad"] = count_include_pad

    def _alias(self):
        return "pool"

    def hybrid_forward(self, F, x):
        channels_last = self._layout is not None and self._layout.endswith("C")
        if channels_last:
            nd_ = len(self._kwargs["kernel"])
            x = F.transpose(x, axes=(0, nd_ + 1) + tuple(range(1, nd_ + 1)))
        y = F.Pooling(x, name="fwd", **self._kwargs)
        if channels_last:
            y = F.transpose(y, axes=(0,) + tuple(range(2, nd_ + 2)) + (1,))
        return y

    def __repr__(self):
        s = "{name}(size={kernel}, stride={stride}, padding={pad}, ceil_mode={ceil_mode})"
        return s.format(name=self.__class__.__name__, ceil_mode=self._kwargs["pooling_convention"] == "full", **self._kwargs)


def _mk_pool(name, nd_, pool_type, global_pool, layout):
    if global_pool:
        def __init__(self, layout=layout, **kwargs):
            _Pooling.__init__(self, (1,) * nd_, None, 0, True, True, pool_type, layout, **kwargs)
    elif pool_type == "max":
        def __init__(self, pool_size=(2,) * nd_, strides=None, padding=0, layout=layout, ceil_mode=False, **kwargs):
            _Pooling.__init__(self, _tup(pool_size, nd_), strides, padding, ceil_mode, False, "max", layout, **kwargs)
    else:
        def __init__(self, pool_size=(2,) * nd_, strides=None, padding=0, ceil_mode=False, layout=layout,
                     count_include_pad=True, **kwargs):
            _Pooling.__init__(self, _tup(pool_size, nd_), strides, padding, ceil_mode, False, "avg", layout,
                              count_include_pad, **kwargs)
    return type(name, (_Pooling,), {"__init__": __init__})


MaxPool1D = _mk_pool("MaxPool1D", 1, "max", False, "NCW")
MaxPool2D = _mk_pool("MaxPool2D", 2, "max", False, "NCHW")
MaxPool3D = _mk_pool("MaxPool3D", 3, "max", False, "NCDHW")
AvgPool1D = _mk_pool("AvgPool1D", 1, "avg", False, "NCW")
AvgPool2D = _mk_pool("AvgPool2D", 2, "avg", False, "NCHW")
AvgPool3D = _mk_pool("AvgPool3D", 3, "avg", False, "NCDHW")
GlobalMaxPool1D = _mk_pool("GlobalMaxPool1D", 1, "max", True, "NCW")
GlobalMaxPool2D = _mk_pool("GlobalMaxPool2D", 2, "max", True, "NCHW")
GlobalMaxPool3D = _mk_pool("GlobalMaxPool3D", 3, "max", True, "NCDHW")
GlobalAvgPool1D = _mk_pool("GlobalAvgPool1D", 1, "avg", True, "NCW")
GlobalAvgPool2D = _mk_pool("GlobalAvgPool2D", 2, "avg", True, "NCHW")
GlobalAvgPool3D = _mk_pool("GlobalAvgPool3D", 3, "avg", True, "NCDHW")


class ReflectionPad2D(HybridBlock):
    def __init__(self, padding=0, **kwargs):
        super(ReflectionPad2D, self).__init__(**kwargs)
        if isinstance(padding, (int, np.integer)):
            padding = (0, 0, 0, 0, padding, padding, padding, padding)
        assert len(padding) == 8
        self._padding = padding

    def hybrid_forward(self, F, x):
        return F.pad(x, mode="reflect", pad_width=self._padding)
