"""AlexNet, VGG, SqueezeNet, DenseNet, MobileNet(v1/v2), Inception-v3
(reference: python/mxnet/gluon/model_zoo/vision/{alexnet,vgg,squeezenet,densenet,mobilenet,inception}.py)."""
from ...block import HybridBlock
from ... import nn

__all__ = ["AlexNet", "alexnet", "VGG", "vgg11", "vgg13", "vgg16", "vgg19", "vgg11_bn", "vgg13_bn", "vgg16_bn",
           "vgg19_bn", "SqueezeNet", "squeezenet1_0", "squeezenet1_1", "DenseNet", "densenet121", "densenet161",
           "densenet169", "densenet201", "MobileNet", "MobileNetV2", "mobilenet1_0", "mobilenet0_75", "mobilenet0_5",
           "mobilenet0_25", "mobilenet_v2_1_0", "mobilenet_v2_0_75", "mobilenet_v2_0_5", "mobilenet_v2_0_25",
           "Inception3", "inception_v3"]


class AlexNet(HybridBlock):
    def __init__(self, classes=1000, **kwargs):
        super(AlexNet, self).__init__(**kwargs)
        with self.name_scope():
            self.features = nn.HybridSequential(prefix="")
            with self.features.name_scope():
                self.features.add(nn.Conv2D(64, kernel_size=11, strides=4, padding=2, activation="relu"))
                self.features.add(nn.MaxPool2D(pool_size=3, strides=2))
                self.features.add(nn.Conv2D(192, kernel_size=5, padding=2, activation="relu"))
                self.features.add(nn.MaxPool2D(pool_size=3, strides=2))
                self.features.add(nn.Conv2D(384, kernel_size=3, padding=1, activation="relu"))
                self.features.add(nn.Conv2D(256, kernel_size=3, padding=1, activation="relu"))
                self.features.add(nn.Conv2D(256, kernel_size=3, padding=1, activation="relu"))
                self.features.add(nn.MaxPool2D(pool_size=3, strides=2))
                self.features.add(nn.Flatten())
                self.features.add(nn.Dense(4096, activation="relu"))
                self.features.add(nn.Dropout(0.5))
                self.features.add(nn.Dense(4096, activation="relu"))
                self.features.add(nn.Dropout(0.5))
            self.output = nn.Dense(classes)

    def hybrid_forward(self, F, x):
        return self.output(self.features(x))


def alexnet(pretrained=False, ctx=None, **kwargs):
    net = AlexNet(**kwargs)
    if ctx is not None:
        net.initialize(ctx=ctx)
    return net


class VGG(HybridBlock):
    def __init__(self, layers, filters, classes=1000, batch_norm=False, **kwargs):
        super(VGG, self).__init__(**kwargs)
        assert len(layers) == len(filters)
        with self.name_scope():
            self.features = nn.HybridSequential(prefix="")
            for i, num in enumerate(layers):
                for _ in range(num):
                    self.features.add(nn.Conv2D(filters[i], kernel_size=3, padding=1))
                    if batch_norm:
                        self.features.add(nn.BatchNorm())
                    self.features.add(nn.Activation("relu"))
                self.features.add(nn.MaxPool2D(strides=2))
            self.features.add(nn.Dense(4096, activation="relu", weight_initializer="normal"))
            self.features.add(nn.Dropout(rate=0.5))
            self.features.add(nn.Dense(4096, activation="relu", weight_initializer="normal"))
            self.features.add(nn.Dropout(rate=0.5))
            self.output = nn.Dense(classes, weight_initializer="normal")

    def hybrid_forward(self, F, x):
        return self.output(self.features(x))


_VGG = {11: ([1, 1, 2, 2, 2], [64, 128, 256, 512, 512]), 13: ([2, 2, 2, 2, 2], [64, 128, 256, 512, 512]),
        16: ([2, 2, 3, 3, 3], [64, 128, 256, 512, 512]), 19: ([2, 2, 4, 4, 4], [64, 128, 256, 512, 512])}


def _vgg(n, bn):
    def f(pretrained=False, ctx=None, **kwargs):
        layers, filters = _VGG[n]
        net = VGG(layers, filters, batch_norm=bn, **kwargs)
        if ctx is not None:
            net.initialize(ctx=ctx)
        return net
    return f


vgg11, vgg13, vgg16, vgg19 = [_vgg(n, False) for n in (11, 13, 16, 19)]
vgg11_bn, vgg13_bn, vgg16_bn, vgg19_bn = [_vgg(n, True) for n in (11, 13, 16, 19)]


class _Fire(HybridBlock):
    def __init__(self, squeeze, e1, e3, **kwargs):
        super(_Fire, self).__init__(**kwargs)
        self.squeeze = nn.Conv2D(squeeze, 1, activation="relu")
        self.e1 = nn.Conv2D(e1, 1, activation="relu")
        self.e3 = nn.Conv2D(e3, 3, padding=1, activation="relu")

    def hybrid_forward(self, F, x):
        x = self.squeeze(x)
        return F.concat(self.e1(x), self.e3(x), dim=1)


class SqueezeNet(HybridBlock):
    def __init__(self, version, classes=1000, **kwargs):
        super(SqueezeNet, self).__init__(**kwargs)
        assert version in ["1.0", "1.1"]
        with self.name_scope():
            self.features = nn.HybridSequential(prefix="")
            if version == "1.0":
                self.features.add(nn.Conv2D(96, kernel_size=7, strides=2, activation="relu"))
                self.features.add(nn.MaxPool2D(pool_size=3, strides=2, ceil_mode=True))
                cfg = [(16, 64, 64), (16, 64, 64), (32, 128, 128), "P", (32, 128, 128), (48, 192, 192), (48, 192, 192),
                       (64, 256, 256), "P", (64, 256, 256)]
            else:
                self.features.add(nn.Conv2D(64, kernel_size=3, strides=2, activation="relu"))
                self.features.add(nn.MaxPool2D(pool_size=3, strides=2, ceil_mode=True))
                cfg = [(16, 64, 64), (16, 64, 64), "P", (32, 128, 128), (32, 128, 128), "P", (48, 192, 192),
                       (48, 192, 192), (64, 256, 256), (64, 256, 256)]
            for c in cfg:
                if c == "P":
                    self.features.add(nn.MaxPool2D(pool_size=3, strides=2, ceil_mode=True))
                else:
                    self.features.add(_Fire(*c))
            self.features.add(nn.Dropout(0.5))
            self.output = nn.HybridSequential(prefix="")
            self.output.add(nn.Conv2D(classes, kernel_size=1, activation="relu"))
            self.output.add(nn.GlobalAvgPool2D())
            self.output.add(nn.Flatten())

    def hybrid_forward(self, F, x):
        return self.output(self.features(x))


def squeezenet1_0(pretrained=False, ctx=None, **kwargs):
    net = SqueezeNet("1.0", **kwargs)
    if ctx is not None:
        net.initialize(ctx=ctx)
    return net


def squeezenet1_1(pretrained=False, ctx=None, **kwargs):
    net = SqueezeNet("1.1", **kwargs)
    if ctx is not None:
        net.initialize(ctx=ctx)
    return net


class _DenseLayer(HybridBlock):
    def __init__(self, growth_rate, bn_size, dropout, **kwargs):
        super(_DenseLayer, self).__init__(**kwargs)
        self.body = nn.HybridSequential(prefix="")
        self.body.add(nn.BatchNorm())
        self.body.add(nn.Activation("relu"))
        self.body.add(nn.Conv2D(bn_size * growth_rate, kernel_size=1, use_bias=False))
        self.body.add(nn.BatchNorm())
        self.body.add(nn.Activation("relu"))
        self.body.add(nn.Conv2D(growth_rate, kernel_size=3, padding=1, use_bias=False))
        if dropout:
            self.body.add(nn.Dropout(dropout))

    def hybrid_forward(self, F, x):
        return F.concat(x, self.body(x), dim=1)


class DenseNet(HybridBlock):
    def __init__(self, num_init_features, growth_rate, block_config, bn_size=4, dropout=0, classes=1000, **kwargs):
        super(DenseNet, self).__init__(**kwargs)
        with self.name_scope():
            self.features = nn.HybridSequential(prefix="")
            self.features.add(nn.Conv2D(num_init_features, kernel_size=7, strides=2, padding=3, use_bias=False))
            self.features.add(nn.BatchNorm())
            self.features.add(nn.Activation("relu"))
            self.features.add(nn.MaxPool2D(pool_size=3, strides=2, padding=1))
            num_features = num_init_features
            for i, num_layers in enumerate(block_config):
                for _ in range(num_layers):
                    self.features.add(_DenseLayer(growth_rate, bn_size, dropout))
                num_features += num_layers * growth_rate
                if i != len(block_config) - 1:
                    self.features.add(nn.BatchNorm())
                    self.features.add(nn.Activation("relu"))
                    self.features.add(nn.Conv2D(num_features // 2, kernel_size=1, use_bias=False))
                    self.features.add(nn.AvgPool2D(pool_size=2, strides=2))
                    num_features //= 2
            self.features.add(nn.BatchNorm())
            self.features.add(nn.Activation("relu"))
            self.features.add(nn.AvgPool2D(pool_size=7))
            self.features.add(nn.Flatten())
            self.output = nn.Dense(classes)

    def hybrid_forward(self, F, x):
        return self.output(self.features(x))


_DENSE = {121: (64, 32, [6, 12, 24, 16]), 161: (96, 48, [6, 12, 36, 24]), 169: (64, 32, [6, 12, 32, 32]),
          201: (64, 32, [6, 12, 48, 32])}


def _densenet(n):
    def f(pretrained=False, ctx=None, **kwargs):
        a, b, c = _DENSE[n]
        net = DenseNet(a, b, c, **kwargs)
        if ctx is not None:
            net.initialize(ctx=ctx)
        return net
    return f


densenet121, densenet161, densenet169, densenet201 = [_densenet(n) for n in (121, 161, 169, 201)]


def _add_conv(out, channels=1, kernel=1, stride=1, pad=0, num_group=1, active=True, relu6=False):
    out.add(nn.Conv2D(channels, kernel, stride, pad, groups=num_group, use_bias=False))
    out.add(nn.BatchNorm(scale=True))
    if active:
        out.add(nn.HybridLambda(lambda F, x: F.clip(x, 0, 6)) if relu6 else nn.Activation("relu"))


class MobileNet(HybridBlock):
    def __init__(self, multiplier=1.0, classes=1000, **kwargs):
        super(MobileNet, self).__init__(**kwargs)
        with self.name_scope():
            self.features = nn.HybridSequential(prefix="")
            with self.features.name_scope():
                _add_conv(self.features, channels=int(32 * multiplier), kernel=3, pad=1, stride=2)
                dw_channels = [int(x * multiplier) for x in [32, 64] + [128] * 2 + [256] * 2 + [512] * 6 + [1024]]
                channels = [int(x * multiplier) for x in [64] + [128] * 2 + [256] * 2 + [512] * 6 + [1024] * 2]
                strides = [1, 2] * 3 + [1] * 5 + [2, 1]
                for dwc, c, s in zip(dw_channels, channels, strides):
                    _add_conv(self.features, channels=dwc, kernel=3, stride=s, pad=1, num_group=dwc)
                    _add_conv(self.features, channels=c)
                self.features.add(nn.GlobalAvgPool2D())
                self.features.add(nn.Flatten())
            self.output = nn.Dense(classes)

    def hybrid_forward(self, F, x):
        return self.output(self.features(x))


class _LinearBottleneck(HybridBlock):
    def __init__(self, in_channels, channels, t, stride, **kwargs):
        super(_LinearBottleneck, self).__init__(**kwargs)
        self.use_shortcut = stride == 1 and in_channels == channels
        with self.name_scope():
            self.out = nn.HybridSequential()
            _add_conv(self.out, in_channels * t, relu6=True)
            _add_conv(self.out, in_channels * t, kernel=3, stride=stride, pad=1, num_group=in_channels * t, relu6=True)
            _add_conv(self.out, channels, active=False, relu6=True)

    def hybrid_forward(self, F, x):
        out = self.out(x)
        return out + x if self.use_shortcut else out


class MobileNetV2(HybridBlock):
    def __init__(self, multiplier=1.0, classes=1000, **kwargs):
        super(MobileNetV2, self).__init__(**kwargs)
        with self.name_scope():
            self.features = nn.HybridSequential(prefix="features_")
            with self.features.name_scope():
                _add_conv(self.features, int(32 * multiplier), kernel=3, stride=2, pad=1, relu6=True)
                in_group = [int(x * multiplier) for x in [32] + [16] + [24] * 2 + [32] * 3 + [64] * 4 + [96] * 3 + [160] * 3]
                c_group = [int(x * multiplier) for x in [16] + [24] * 2 + [32] * 3 + [64] * 4 + [96] * 3 + [160] * 3 + [320]]
                ts = [1] + [6] * 16
                strides = [1, 2] * 2 + [1, 1, 2] + [1] * 6 + [2] + [1] * 3
                for in_c, c, t, s in zip(in_group, c_group, ts, strides):
                    self.features.add(_LinearBottleneck(in_channels=in_c, channels=c, t=t, stride=s))
                last = int(1280 * multiplier) if multiplier > 1.0 else 1280
                _add_conv(self.features, last, relu6=True)
                self.features.add(nn.GlobalAvgPool2D())
            self.output = nn.HybridSequential(prefix="output_")
            with self.output.name_scope():
                self.output.add(nn.Conv2D(classes, 1, use_bias=False, prefix="pred_"), nn.Flatten())

    def hybrid_forward(self, F, x):
        return self.output(self.features(x))


def _mobilenet(cls, mult):
    def f(pretrained=False, ctx=None, **kwargs):
        net = cls(mult, **kwargs)
        if ctx is not None:
            net.initialize(ctx=ctx)
        return net
    return f


mobilenet1_0, mobilenet0_75, mobilenet0_5, mobilenet0_25 = [_mobilenet(MobileNet, m) for m in (1.0, 0.75, 0.5, 0.25)]
mobilenet_v2_1_0, mobilenet_v2_0_75, mobilenet_v2_0_5, mobilenet_v2_0_25 = \
    [_mobilenet(MobileNetV2, m) for m in (1.0, 0.75, 0.5, 0.25)]


def _bconv(channels, kernel, stride=1, pad=0):
    out = nn.HybridSequential(prefix="")
    out.add(nn.Conv2D(channels, kernel, stride, pad, use_bias=False))
    out.add(nn.BatchNorm(epsilon=0.001))
    out.add(nn.Activation("relu"))
    return out


class _Branches(HybridBlock):
    """Run each branch on the same input and concatenate along channels."""

    def __init__(self, branches, **kwargs):
        super(_Branches, self).__init__(**kwargs)
        for i, b in enumerate(branches):
            self.register_child(b, "b%d" % i)

    def hybrid_forward(self, F, x):
        return F.concat(*[b(x) for b in self._children.values()], dim=1)


def _seq(*blocks):
    s = nn.HybridSequential(prefix="")
    s.add(*blocks)
    return s


def _incA(pool_features):
    return _Branches([_bconv(64, 1), _seq(_bconv(48, 1), _bconv(64, 5, pad=2)),
                      _seq(_bconv(64, 1), _bconv(96, 3, pad=1), _bconv(96, 3, pad=1)),
                      _seq(nn.AvgPool2D(3, 1, 1), _bconv(pool_features, 1))])


def _incB():
    return _Branches([_bconv(384, 3, 2), _seq(_bconv(64, 1), _bconv(96, 3, pad=1), _bconv(96, 3, 2)), nn.MaxPool2D(3, 2)])


def _incC(c7):
    return _Branches([_bconv(192, 1), _seq(_bconv(c7, 1), _bconv(c7, (1, 7), pad=(0, 3)), _bconv(192, (7, 1), pad=(3, 0))),
                      _seq(_bconv(c7, 1), _bconv(c7, (7, 1), pad=(3, 0)), _bconv(c7, (1, 7), pad=(0, 3)),
                           _bconv(c7, (7, 1), pad=(3, 0)), _bconv(192, (1, 7), pad=(0, 3))),
                      _seq(nn.AvgPool2D(3, 1, 1), _bconv(192, 1))])


def _incD():
    return _Branches([_seq(_bconv(192, 1), _bconv(320, 3, 2)),
                      _seq(_bconv(192, 1), _bconv(192, (1, 7), pad=(0, 3)), _bconv(192, (7, 1), pad=(3, 0)), _bconv(192, 3, 2)),
                      nn.MaxPool2D(3, 2)])


def _incE():
    return _Branches([_bconv(320, 1),
                      _seq(_bconv(384, 1), _Branches([_bconv(384, (1, 3), pad=(0, 1)), _bconv(384, (3, 1), pad=(1, 0))])),
                      _seq(_bconv(448, 1), _bconv(384, 3, pad=1),
                           _Branches([_bconv(384, (1, 3), pad=(0, 1)), _bconv(384, (3, 1), pad=(1, 0))])),
                      _seq(nn.AvgPool2D(3, 1, 1), _bconv(192, 1))])


class Inception3(HybridBlock):
    def __init__(self, classes=1000, **kwargs):
        super(Inception3, self).__init__(**kwargs)
        with self.name_scope():
            self.features = nn.HybridSequential(prefix="")
            self.features.add(_bconv(32, 3, 2), _bconv(32, 3), _bconv(64, 3, pad=1), nn.MaxPool2D(3, 2),
                              _bconv(80, 1), _bconv(192, 3), nn.MaxPool2D(3, 2),
                              _incA(32), _incA(64), _incA(64), _incB(), _incC(128), _incC(160), _incC(160), _incC(192),
                              _incD(), _incE(), _incE(), nn.AvgPool2D(8), nn.Dropout(0.5))
            self.output = nn.Dense(classes)

    def hybrid_forward(self, F, x):
        return self.output(self.features(x))


def inception_v3(pretrained=False, ctx=None, **kwargs):
    net = Inception3(**kwargs)
    if ctx is not None:
        net.initialize(ctx=ctx)
    return net
