"""ResNet (v1b bottleneck / basic) in channels-last bf16, built on the framework's own kernels.

Reference models: python/mxnet/gluon/model_zoo/vision/resnet.py:231-529 (resnet{18..152}_v1/v2) and the
GluonCV ``resnet50_v1b`` used by example/dynamic-training/train_resnet.py (stride on the 3x3 conv).
The reference runs each of Conv / BatchNorm / Activation / add as a separate engine op through cuDNN.
Here, per bottleneck:

  1x1 conv  -> tcgen05 GEMM over pixels (ops.functional.conv1x1_nhwc), fwd/dgrad/wgrad all ours
  BN+ReLU   -> one stats pass + one apply pass (ops.functional.bn_act)
  BN+add+ReLU at the block end -> the same two passes (residual and ReLU fused in)
  3x3 / 7x7 conv -> cuDNN via torch on the channels-last view (library op; replaced by an implicit-GEMM
                    tcgen05 kernel when available)

Activations stay [N, H, W, C] contiguous; weights of 1x1 convs are [Cout, Cin] matrices, spatial kernels
are stored [Cout, kh, kw, Cin] (KRSC) and exposed to cuDNN as channels_last (Cout, Cin, kh, kw) views.
This is the *functional core*; ``mxb200.gluon.model_zoo.vision`` wraps it as HybridBlocks.
"""
import math

import torch
import torch.nn.functional as F

from ..ops import functional as K

__all__ = ["ResNetSpec", "resnet_spec", "init_resnet_params", "resnet_forward", "RESNET_SPECS"]

RESNET_SPECS = {
    18: ("basic", [2, 2, 2, 2], [64, 64, 128, 256, 512]),
    34: ("basic", [3, 4, 6, 3], [64, 64, 128, 256, 512]),
    50: ("bottleneck", [3, 4, 6, 3], [64, 256, 512, 1024, 2048]),
    101: ("bottleneck", [3, 4, 23, 3], [64, 256, 512, 1024, 2048]),
    152: ("bottleneck", [3, 8, 36, 3], [64, 256, 512, 1024, 2048]),
}


class ResNetSpec(object):
    def __init__(self, num_layers=50, classes=1000, version="v1b", in_channels=3, thumbnail=False):
        self.block, self.layers, self.channels = RESNET_SPECS[num_layers]
        self.num_layers, self.classes, self.version = num_layers, classes, version
        self.in_channels, self.thumbnail = in_channels, thumbnail


def resnet_spec(num_layers=50, **kw):
    return ResNetSpec(num_layers, **kw)


def _conv_w(shape_oihw, dtype, device, gen):
    """MSRA (He) init, stored KRSC; returns the (O,I,kh,kw) channels_last view (1x1 -> [O, I] matrix)."""
    o, i, kh, kw = shape_oihw
    std = math.sqrt(2.0 / (i * kh * kw))
    w = torch.randn((o, kh, kw, i), generator=gen, dtype=torch.float32) * std
    w = w.to(device=device, dtype=dtype)
    if kh == 1 and kw == 1:
        return w.reshape(o, i)
    return w.permute(0, 3, 1, 2)


def init_resnet_params(spec, dtype=torch.bfloat16, device="cpu", seed=0):
    """Returns (params, aux): ordered dicts name -> tensor.  BN gamma/beta/stats are fp32 (as in the
    reference's fp16 recipe, where BatchNorm parameters stay fp32)."""
    gen = torch.Generator().manual_seed(seed)
    P, A = {}, {}

    def bn(name, c):
        P[name + "_gamma"] = torch.ones(c, dtype=torch.float32, device=device)
        P[name + "_beta"] = torch.zeros(c, dtype=torch.float32, device=device)
        A[name + "_running_mean"] = torch.zeros(c, dtype=torch.float32, device=device)
        A[name + "_running_var"] = torch.ones(c, dtype=torch.float32, device=device)

    c0 = spec.channels[0]
    if spec.thumbnail:
        P["conv0_weight"] = _conv_w((c0, spec.in_channels, 3, 3), dtype, device, gen)
    else:
        P["conv0_weight"] = _conv_w((c0, spec.in_channels, 7, 7), dtype, device, gen)
    bn("bn0", c0)
    cin = c0
    for s, (n, cout) in enumerate(zip(spec.layers, spec.channels[1:])):
        for b in range(n):
            pre = "stage%d_unit%d_" % (s + 1, b + 1)
            stride = 2 if (b == 0 and s > 0) else 1
            if spec.block == "bottleneck":
                mid = cout // 4
                P[pre + "conv1_weight"] = _conv_w((mid, cin, 1, 1), dtype, device, gen); bn(pre + "bn1", mid)
                P[pre + "conv2_weight"] = _conv_w((mid, mid, 3, 3), dtype, device, gen); bn(pre + "bn2", mid)
                P[pre + "conv3_weight"] = _conv_w((cout, mid, 1, 1), dtype, device, gen); bn(pre + "bn3", cout)
            else:
                P[pre + "conv1_weight"] = _conv_w((cout, cin, 3, 3), dtype, device, gen); bn(pre + "bn1", cout)
                P[pre + "conv2_weight"] = _conv_w((cout, cout, 3, 3), dtype, device, gen); bn(pre + "bn2", cout)
            if stride != 1 or cin != cout:
                P[pre + "down_weight"] = _conv_w((cout, cin, 1, 1), dtype, device, gen); bn(pre + "down_bn", cout)
            cin = cout
    bound = 1.0 / math.sqrt(cin)
    P["fc_weight"] = ((torch.rand((spec.classes, cin), generator=gen) * 2 - 1) * bound).to(device=device, dtype=dtype)
    P["fc_bias"] = torch.zeros(spec.classes, dtype=torch.float32, device=device)
    return P, A


def _conv_spatial(x, w, stride, pad):
    """kxk convolution on an NHWC tensor through the channels_last NCHW view (cuDNN)."""
    y = F.conv2d(x.permute(0, 3, 1, 2), w, None, stride, pad)
    return y.permute(0, 2, 3, 1)


def _bn(x, P, A, name, relu, residual=None, training=True, eps=1e-5, momentum=0.9):
    return K.bn_act(x, P[name + "_gamma"], P[name + "_beta"], A[name + "_running_mean"], A[name + "_running_var"],
                    residual=residual, relu=relu, eps=eps, momentum=momentum, training=training)


def resnet_forward(spec, P, A, x, training=True, eps=1e-5, momentum=0.9):
    """x: [N, H, W, C] channels-last.  Returns logits [N, classes] (bf16)."""
    if spec.thumbnail:
        y = _conv_spatial(x, P["conv0_weight"], 1, 1)
        y = _bn(y.contiguous(), P, A, "bn0", True, None, training, eps, momentum)
    else:
        y = _conv_spatial(x, P["conv0_weight"], 2, 3)
        y = _bn(y.contiguous(), P, A, "bn0", True, None, training, eps, momentum)
        y = F.max_pool2d(y.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    cin = spec.channels[0]
    for s, (n, cout) in enumerate(zip(spec.layers, spec.channels[1:])):
        for b in range(n):
            pre = "stage%d_unit%d_" % (s + 1, b + 1)
            stride = 2 if (b == 0 and s > 0) else 1
            res = y
            if (pre + "down_weight") in P:
                res = K.conv1x1_nhwc(y, P[pre + "down_weight"], stride)
                res = _bn(res, P, A, pre + "down_bn", False, None, training, eps, momentum)
            if spec.block == "bottleneck":
                # v1b: stride lives on the 3x3 (v1 puts it on the first 1x1)
                s1, s2 = (stride, 1) if spec.version == "v1" else (1, stride)
                o = K.conv1x1_nhwc(y, P[pre + "conv1_weight"], s1)
                o = _bn(o, P, A, pre + "bn1", True, None, training, eps, momentum)
                o = _conv_spatial(o, P[pre + "conv2_weight"], s2, 1)
                o = _bn(o.contiguous(), P, A, pre + "bn2", True, None, training, eps, momentum)
                o = K.conv1x1_nhwc(o, P[pre + "conv3_weight"], 1)
                y = _bn(o, P, A, pre + "bn3", True, res, training, eps, momentum)
            else:
                o = _conv_spatial(y, P[pre + "conv1_weight"], stride, 1)
                o = _bn(o.contiguous(), P, A, pre + "bn1", True, None, training, eps, momentum)
                o = _conv_spatial(o, P[pre + "conv2_weight"], 1, 1)
                y = _bn(o.contiguous(), P, A, pre + "bn2", True, res, training, eps, momentum)
            cin = cout
    pooled = y.float().mean(dim=(1, 2)).to(y.dtype)  # global average pool
    return K.linear(pooled, P["fc_weight"], P["fc_bias"])
