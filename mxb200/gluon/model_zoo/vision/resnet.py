"""ResNets (reference: python/mxnet/gluon/model_zoo/vision/resnet.py:36-529).

Two families:

* ``ResNetV1`` / ``ResNetV2`` (+ BasicBlockV1/V2, BottleneckV1/V2): the reference's layer-by-layer Gluon
  definitions in NCHW, composed from nn.Conv2D / nn.BatchNorm / nn.Activation.
* ``FusedResNet`` (``get_resnet(..., layout='NHWC')`` and ``resnet50_v1b``): the channels-last bf16
  production model.  Its parameters are ordinary Gluon Parameters (save/load/Trainer work unchanged) but
  ``hybrid_forward`` is one call into ``mxb200.models.resnet.resnet_forward``, i.e. tcgen05 GEMMs for the
  1x1 convolutions and fused BN(+add)+ReLU kernels -- the analogue of a hybridized CachedOp graph.

Provenance: adapted from Apache MXNet (incubating) 1.3.1 ``python/mxnet/gluon/model_zoo/vision/resnet.py`` (Apache License 2.0).  The public API, argument
names and control flow follow the original so that user scripts keep working; the tensor maths runs on torch /
the mxb200 kernels.  This file is API-compatibility glue, not original B200 work ("Provenance").
"""
import numpy as np
import torch

from ...block import HybridBlock
from ... import nn
from ....context import cpu
from ....ndarray import NDArray
from ....ndarray.ndarray import _op
from .... import autograd
from ....models import resnet as _R

__all__ = ["resnet200_v1", "resnet269_v1", "resnet200_v2", "resnet269_v2", "ResNetV1", "ResNetV2", "BasicBlockV1", "BasicBlockV2", "BottleneckV1", "BottleneckV2", "FusedResNet",
           "resnet18_v1", "resnet34_v1", "resnet50_v1", "resnet101_v1", "resnet152_v1", "resnet18_v2", "resnet34_v2",
           "resnet50_v2", "resnet101_v2", "resnet152_v2", "resnet50_v1b", "resnet101_v1b", "resnet152_v1b",
           "get_resnet"]


def _conv3x3(channels, stride, in_channels):
    return nn.Conv2D(channels, kernel_size=3, strides=stride, padding=1, use_bias=False, in_channels=in_channels)


class BasicBlockV1(HybridBlock):
    def __init__(self, channels, stride, downsample=False, in_channels=0, **kwargs):
        super(BasicBlockV1, self).__init__(**kwargs)
        self.body = nn.HybridSequential(prefix="")
        self.body.add(_conv3x3(channels, stride, in_channels))
        self.body.add(nn.BatchNorm())
        self.body.add(nn.Activation("relu"))
        self.body.add(_conv3x3(channels, 1, channels))
        self.body.add(nn.BatchNorm())
        if downsample:
            self.downsample = nn.HybridSequential(prefix="")
            self.downsample.add(nn.Conv2D(channels, kernel_size=1, strides=stride, use_bias=False, in_channels=in_channels))
            self.downsample.add(nn.BatchNorm())
        else:
            self.downsample = None

    def hybrid_forward(self, F, x):
        residual = x
        x = self.body(x)
        if self.downsample:
            residual = self.downsample(residual)
        return F.Activation(residual + x, act_type="relu")


class BottleneckV1(HybridBlock):
    def __init__(self, channels, stride, downsample=False, in_channels=0, **kwargs):
        super(BottleneckV1, self).__init__(**kwargs)
        self.body = nn.HybridSequential(prefix="")
        self.body.add(nn.Conv2D(channels // 4, kernel_size=1, strides=stride))
        self.body.add(nn.BatchNorm())
        self.body.add(nn.Activation("relu"))
        self.body.add(_conv3x3(channels // 4, 1, channels // 4))
        self.body.add(nn.BatchNorm())
        self.body.add(nn.Activation("relu"))
        self.body.add(nn.Conv2D(channels, kernel_size=1, strides=1))
        self.body.add(nn.BatchNorm())
        if downsample:
            self.downsample = nn.HybridSequential(prefix="")
            self.downsample.add(nn.Conv2D(channels, kernel_size=1, strides=stride, use_bias=False, in_channels=in_channels))
            self.downsample.add(nn.BatchNorm())
        else:
            self.downsample = None

    def hybrid_forward(self, F, x):
        residual = x
        x = self.body(x)
        if self.downsample:
            residual = self.downsample(residual)
        return F.Activation(x + residual, act_type="relu")


class BasicBlockV2(HybridBlock):
    def __init__(self, channels, stride, downsample=False, in_channels=0, **kwargs):
        super(BasicBlockV2, self).__init__(**kwargs)
        self.bn1 = nn.BatchNorm()
        self.conv1 = _conv3x3(channels, stride, in_channels)
        self.bn2 = nn.BatchNorm()
        self.conv2 = _conv3x3(channels, 1, channels)
        self.downsample = nn.Conv2D(channels, 1, stride, use_bias=False, in_channels=in_channels) if downsample else None

    def hybrid_forward(self, F, x):
        residual = x
        x = self.bn1(x)
        x = F.Activation(x, act_type="relu")
        if self.downsample:
            residual = self.downsample(x)
        x = self.conv1(x)
        x = self.bn2(x)
        x = F.Activation(x, act_type="relu")
        x = self.conv2(x)
        return x + residual


class BottleneckV2(HybridBlock):
    def __init__(self, channels, stride, downsample=False, in_channels=0, **kwargs):
        super(BottleneckV2, self).__init__(**kwargs)
        self.bn1 = nn.BatchNorm()
        self.conv1 = nn.Conv2D(channels // 4, kernel_size=1, strides=1, use_bias=False)
        self.bn2 = nn.BatchNorm()
        self.conv2 = _conv3x3(channels // 4, stride, channels // 4)
        self.bn3 = nn.BatchNorm()
        self.conv3 = nn.Conv2D(channels, kernel_size=1, strides=1, use_bias=False)
        self.downsample = nn.Conv2D(channels, 1, stride, use_bias=False, in_channels=in_channels) if downsample else None

    def hybrid_forward(self, F, x):
        residual = x
        x = self.bn1(x)
        x = F.Activation(x, act_type="relu")
        if self.downsample:
            residual = self.downsample(x)
        x = self.conv1(x)
        x = self.bn2(x)
        x = F.Activation(x, act_type="relu")
        x = self.conv2(x)
        x = self.bn3(x)
        x = F.Activation(x, act_type="relu")
        x = self.conv3(x)
        return x + residual


class ResNetV1(HybridBlock):
    def __init__(self, block, layers, channels, classes=1000, thumbnail=False, **kwargs):
        super(ResNetV1, self).__init__(**kwargs)
        assert len(layers) == len(channels) - 1
        with self.name_scope():
            self.features = nn.HybridSequential(prefix="")
            if thumbnail:
                self.features.add(_conv3x3(channels[0], 1, 0))
            else:
                self.features.add(nn.Conv2D(channels[0], 7, 2, 3, use_bias=False))
                self.features.add(nn.BatchNorm())
                self.features.add(nn.Activation("relu"))
                self.features.add(nn.MaxPool2D(3, 2, 1))
            for i, num_layer in enumerate(layers):
                stride = 1 if i == 0 else 2
                self.features.add(self._make_layer(block, num_layer, channels[i + 1], stride, i + 1,
                                                   in_channels=channels[i]))
            self.features.add(nn.GlobalAvgPool2D())
            self.output = nn.Dense(classes, in_units=channels[-1])

    def _make_layer(self, block, layers, channels, stride, stage_index, in_channels=0):
        layer = nn.HybridSequential(prefix="stage%d_" % stage_index)
        with layer.name_scope():
            layer.add(block(channels, stride, channels != in_channels, in_channels=in_channels, prefix=""))
            for _ in range(layers - 1):
                layer.add(block(channels, 1, False, in_channels=channels, prefix=""))
        return layer

    def hybrid_forward(self, F, x):
        x = self.features(x)
        return self.output(x)


class ResNetV2(HybridBlock):
    def __init__(self, block, layers, channels, classes=1000, thumbnail=False, **kwargs):
        super(ResNetV2, self).__init__(**kwargs)
        assert len(layers) == len(channels) - 1
        with self.name_scope():
            self.features = nn.HybridSequential(prefix="")
            self.features.add(nn.BatchNorm(scale=False, center=False))
            if thumbnail:
                self.features.add(_conv3x3(channels[0], 1, 0))
            else:
                self.features.add(nn.Conv2D(channels[0], 7, 2, 3, use_bias=False))
                self.features.add(nn.BatchNorm())
                self.features.add(nn.Activation("relu"))
                self.features.add(nn.MaxPool2D(3, 2, 1))
            in_channels = channels[0]
            for i, num_layer in enumerate(layers):
                stride = 1 if i == 0 else 2
                self.features.add(self._make_layer(block, num_layer, channels[i + 1], stride, i + 1,
                                                   in_channels=in_channels))
                in_channels = channels[i + 1]
            self.features.add(nn.BatchNorm())
            self.features.add(nn.Activation("relu"))
            self.features.add(nn.GlobalAvgPool2D())
            self.features.add(nn.Flatten())
            self.output = nn.Dense(classes, in_units=in_channels)

    _make_layer = ResNetV1._make_layer

    def hybrid_forward(self, F, x):
        x = self.features(x)
        return self.output(x)


class FusedResNet(HybridBlock):
    """Channels-last bf16 ResNet running on the framework's own kernels (see module docstring).

    Input: NDArray [N, H, W, C] (``layout='NHWC'``) or [N, C, H, W] (``layout='NCHW'``, permuted on entry).
    """

    def __init__(self, num_layers=50, classes=1000, version="v1b", thumbnail=False, dtype="bfloat16", layout="NHWC",
                 epsilon=1e-5, momentum=0.9, seed=0, **kwargs):
        super(FusedResNet, self).__init__(**kwargs)
        from ....base import torch_dtype
        self._spec = _R.ResNetSpec(num_layers, classes=classes, version=version, thumbnail=thumbnail)
        self._layout = layout
        self._eps, self._momentum = epsilon, momentum
        self._tdtype = torch_dtype(dtype)
        P, A = _R.init_resnet_params(self._spec, dtype=self._tdtype, device="cpu", seed=seed)
        self._pnames, self._anames = list(P.keys()), list(A.keys())
        self._init_values = {}
        with self.name_scope():
            for n, t in P.items():
                p = self.params.get(n, shape=tuple(t.shape), dtype=dtype if t.dtype is self._tdtype else "float32")
                self._init_values[p.name] = t
                self._reg_params[n] = p
            for n, t in A.items():
                p = self.params.get(n, shape=tuple(t.shape), dtype="float32", grad_req="null", differentiable=False)
                self._init_values[p.name] = t
                self._reg_params[n] = p

    def initialize(self, init=None, ctx=None, verbose=False, force_reinit=False):
        """MSRA-initialised weights (the values drawn at construction); ``init`` is accepted for API parity."""
        from ....context import Context, current_context
        if ctx is None:
            ctx = [current_context()]
        if isinstance(ctx, Context):
            ctx = [ctx]
        for name, p in self.params.items():
            if p._data is not None and not force_reinit:
                continue
            t = self._init_values[name]
            with autograd.pause():
                data = []
                for c in ctx:
                    d = t.to(c.torch_device)
                    if t.dim() == 4:
                        # keep KRSC physical layout on the device copy
                        d = d.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
                    data.append(NDArray(d.clone(memory_format=torch.preserve_format), c))
                p._ctx_list = list(ctx)
                p._data = data
                p._deferred_init = ()
                p._init_grad()

    def cast(self, dtype):
        pass  # dtype is fixed at construction (weights bf16, BN parameters/statistics fp32)

    def hybrid_forward(self, F, x, **params):
        if not isinstance(x, NDArray):
            raise NotImplementedError("FusedResNet has no symbolic graph; use ResNetV1/ResNetV2 for export()")
        training = autograd.is_training()
        P = {n: params[n]._t for n in self._pnames}
        A = {n: params[n]._t for n in self._anames}
        G = {n: self._reg_params[n]._grad_sink for n in self._pnames if self._reg_params[n]._grad_sink is not None}
        if not G or not autograd.is_recording():
            G = None
        spec, eps, mom = self._spec, self._eps, self._momentum
        nchw = self._layout == "NCHW"

        def run(xt):
            if nchw:
                xt = xt.permute(0, 2, 3, 1)
            if xt.dtype is not self._tdtype:
                xt = xt.to(self._tdtype)
            if not xt.is_contiguous():
                xt = xt.contiguous()
            return _R.resnet_forward(spec, P, A, xt, training=training, eps=eps, momentum=mom, G=G)
        return _op(run, x)


_SPEC = {18: ("basic_block", [2, 2, 2, 2], [64, 64, 128, 256, 512]),
         34: ("basic_block", [3, 4, 6, 3], [64, 64, 128, 256, 512]),
         50: ("bottle_neck", [3, 4, 6, 3], [64, 256, 512, 1024, 2048]),
         101: ("bottle_neck", [3, 4, 23, 3], [64, 256, 512, 1024, 2048]),
         152: ("bottle_neck", [3, 8, 36, 3], [64, 256, 512, 1024, 2048]),
         # 200 / 269 layers: the deep ImageNet variants of the reference's symbol library
         # (example/image-classification/symbols/resnet.py:171-179), used by tools/bandwidth/measure.py
         200: ("bottle_neck", [3, 24, 36, 3], [64, 256, 512, 1024, 2048]),
         269: ("bottle_neck", [3, 30, 48, 8], [64, 256, 512, 1024, 2048])}
_NETS = [ResNetV1, ResNetV2]
_BLOCKS = [{"basic_block": BasicBlockV1, "bottle_neck": BottleneckV1},
           {"basic_block": BasicBlockV2, "bottle_neck": BottleneckV2}]


def get_resnet(version, num_layers, pretrained=False, ctx=None, root=None, layout=None, dtype=None, **kwargs):
    """``version`` 1 / 2 -> reference Gluon definitions; ``'v1b'`` (or layout='NHWC') -> FusedResNet."""
    assert num_layers in _SPEC, "Invalid number of layers: %d. Options are %s" % (num_layers, str(_SPEC.keys()))
    pfile = None
    if pretrained:
        from ..model_store import get_model_file
        pfile = get_model_file("resnet%d_v%s" % (num_layers, version), root=root)    # local model store, nothing is downloaded
    if version in ("v1b", "1b") or layout == "NHWC":
        ver = "v1b" if version in ("v1b", "1b", 1) else "v1"
        net = FusedResNet(num_layers, version=ver, dtype=dtype or "bfloat16", layout=layout or "NHWC", **kwargs)
        if ctx is not None:
            net.initialize(ctx=ctx)
        if pfile is not None:
            net.load_parameters(pfile, ctx=ctx)
        return net
    block_type, layers, channels = _SPEC[num_layers]
    assert 1 <= version <= 2, "Invalid resnet version: %d. Options are 1 and 2." % version
    net = _NETS[version - 1](_BLOCKS[version - 1][block_type], layers, channels, **kwargs)
    if ctx is not None:
        net.initialize(ctx=ctx)
    if pfile is not None:
        net.load_parameters(pfile, ctx=ctx)
    if dtype is not None:
        net.cast(dtype)
    return net


def _mk(version, n):
    def f(**kwargs):
        return get_resnet(version, n, **kwargs)
    return f


resnet18_v1, resnet34_v1, resnet50_v1, resnet101_v1, resnet152_v1 = [_mk(1, n) for n in (18, 34, 50, 101, 152)]
resnet18_v2, resnet34_v2, resnet50_v2, resnet101_v2, resnet152_v2 = [_mk(2, n) for n in (18, 34, 50, 101, 152)]
resnet50_v1b, resnet101_v1b, resnet152_v1b = [_mk("v1b", n) for n in (50, 101, 152)]
resnet200_v1, resnet269_v1, resnet200_v2, resnet269_v2 = _mk(1, 200), _mk(1, 269), _mk(2, 200), _mk(2, 269)
