"""ResNet (v1b bottleneck / basic) in channels-last bf16, built on the framework's own kernels.

Reference models: python/mxnet/gluon/model_zoo/vision/resnet.py:231-529 (resnet{18..152}_v1/v2) and the
GluonCV ``resnet50_v1b`` used by example/dynamic-training/train_resnet.py (stride on the 3x3 conv).
The reference runs each of Conv / BatchNorm / Activation / add as a separate engine op through cuDNN.
Here, per bottleneck:

  1x1 conv  -> tcgen05 GEMM over pixels (ops.functional.conv1x1_nhwc), fwd/dgrad/wgrad all ours
  BN+ReLU   -> one stats pass + one apply pass (ops.functional.bn_act)
  BN+add+ReLU at the block end -> the same two passes (residual and ReLU fused in)
  3x3 conv  -> implicit-GEMM tcgen05 kernel (im2col-mode TMA feeds the same GEMM pipeline; BN statistics in the
               epilogue; fwd / dgrad / wgrad; csrc/cuda/gemm_tcgen05.cu CONV modes)
  7x7 stem  -> small-channel im2col kernel + tcgen05 GEMM (ops.functional.stem_conv_nhwc)

Activations stay [N, H, W, C] contiguous; weights of 1x1 convs are [Cout, Cin] matrices, spatial kernels
are stored [Cout, kh, kw, Cin] (KRSC) and exposed to cuDNN as channels_last (Cout, Cin, kh, kw) views.
This is the *functional core*; ``mxb200.gluon.model_zoo.vision`` wraps it as HybridBlocks.
"""
import math

import weakref

import torch
import torch.nn.functional as F

from ..ops import functional as K
from ..base import getenv

__all__ = ["ResNetSpec", "resnet_spec", "resnet_param_spec", "init_resnet_params", "resnet_forward", "RESNET_SPECS"]

RESNET_SPECS = {
    18: ("basic", [2, 2, 2, 2], [64, 64, 128, 256, 512]),
    34: ("basic", [3, 4, 6, 3], [64, 64, 128, 256, 512]),
    50: ("bottleneck", [3, 4, 6, 3], [64, 256, 512, 1024, 2048]),
    101: ("bottleneck", [3, 4, 23, 3], [64, 256, 512, 1024, 2048]),
    152: ("bottleneck", [3, 8, 36, 3], [64, 256, 512, 1024, 2048]),
    200: ("bottleneck", [3, 24, 36, 3], [64, 256, 512, 1024, 2048]),
}


class ResNetSpec(object):
    def __init__(self, num_layers=50, classes=1000, version="v1b", in_channels=3, thumbnail=False):
        self.block, self.layers, self.channels = RESNET_SPECS[num_layers]
        self.num_layers, self.classes, self.version = num_layers, classes, version
        self.in_channels, self.thumbnail = in_channels, thumbnail


def resnet_spec(num_layers=50, **kw):
    return ResNetSpec(num_layers, **kw)


def resnet_param_spec(spec):
    """Ordered list of (name, shape, dtype_kind, is_aux) for every parameter / auxiliary state of the fused ResNet
    (dtype_kind: 'w' = network dtype, 'f' = always fp32).  1x1 convolution weights are 2-D [O, I] (they are GEMM operands),
    k x k weights are (O, I, kh, kw)."""
    out = []

    def conv(name, o, i, k):
        out.append((name, (o, i) if k == 1 else (o, i, k, k), "w", False))

    def bn(name, c):
        out.extend([(name + "_gamma", (c,), "f", False), (name + "_beta", (c,), "f", False),
                    (name + "_running_mean", (c,), "f", True), (name + "_running_var", (c,), "f", True)])
    c0 = spec.channels[0]
    conv("conv0_weight", c0, spec.in_channels, 3 if spec.thumbnail else 7)
    bn("bn0", c0)
    cin = c0
    for s, (n, cout) in enumerate(zip(spec.layers, spec.channels[1:])):
        for b in range(n):
            pre = "stage%d_unit%d_" % (s + 1, b + 1)
            stride = 2 if (b == 0 and s > 0) else 1
            if spec.block == "bottleneck":
                mid = cout // 4
                conv(pre + "conv1_weight", mid, cin, 1); bn(pre + "bn1", mid)
                conv(pre + "conv2_weight", mid, mid, 3); bn(pre + "bn2", mid)
                conv(pre + "conv3_weight", cout, mid, 1); bn(pre + "bn3", cout)
            else:
                conv(pre + "conv1_weight", cout, cin, 3); bn(pre + "bn1", cout)
                conv(pre + "conv2_weight", cout, cout, 3); bn(pre + "bn2", cout)
            if stride != 1 or cin != cout:
                conv(pre + "down_weight", cout, cin, 1); bn(pre + "down_bn", cout)
            cin = cout
    out.append(("fc_weight", (spec.classes, cin), "w", False))
    out.append(("fc_bias", (spec.classes,), "f", False))
    return out


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


def _stem_conv(x, w, stride, pad, pad_to):
    """7x7 (or 3x3) stem convolution.  With 3 input channels cuDNN falls back to slow generic kernels
    (1.4 ms fwd + 0.75 ms wgrad at batch 256 on B200); zero-padding the input/weight channels to 8 puts the stem
    on the tensor-core implicit-GEMM path.  The parameter keeps its (O,3,k,k) shape: padding is applied on the fly."""
    c = x.shape[-1]
    if pad_to and c < pad_to and x.is_cuda:
        x = F.pad(x, (0, pad_to - c))
        w = F.pad(w, (0, 0, 0, 0, 0, pad_to - c)).contiguous(memory_format=torch.channels_last)
    return _conv_spatial(x, w, stride, pad)


def _sink(G, name):
    return G.get(name) if G is not None else None


def _conv_bn(x, P, A, G, wname, bnname, relu, stride, pad, residual=None, training=True, eps=1e-5, momentum=0.9):
    """k x k convolution + BatchNorm(+residual)(+ReLU).  On B200 the convolution is the implicit-GEMM tcgen05 kernel
    (im2col-mode TMA), its epilogue emits the BatchNorm statistics, dW lands in the gradient arena; otherwise (CPU, odd
    channel counts) the library convolution followed by the stand-alone statistics pass."""
    w = P[wname]
    if K.conv_nhwc_ok(x, w):
        ws = K.bn_stats_workspace_for(x.device, w.shape[0]) if training else None
        o = K.conv_nhwc(x, w, stride, pad, stats=ws, sink_w=_sink(G, wname))
        return _bn(o, P, A, G, bnname, relu, residual, training, eps, momentum, pre_stats=ws)
    o = _conv_spatial(x, w, stride, pad)
    return _bn(o.contiguous(), P, A, G, bnname, relu, residual, training, eps, momentum)


def _bn(x, P, A, G, name, relu, residual=None, training=True, eps=1e-5, momentum=0.9, pre_stats=None):
    return K.bn_act(x, P[name + "_gamma"], P[name + "_beta"], A[name + "_running_mean"], A[name + "_running_var"],
                    residual=residual, relu=relu, eps=eps, momentum=momentum, training=training, pre_stats=pre_stats,
                    sink_g=_sink(G, name + "_gamma"), sink_b=_sink(G, name + "_beta"))


def _conv1x1_bn(x, P, A, G, wname, bnname, relu, stride=1, residual=None, training=True, eps=1e-5, momentum=0.9,
                passthrough=False):
    """1x1 conv (tcgen05 GEMM) whose epilogue also produces the BatchNorm statistics of its output, followed by
    the fused BN(+residual)(+ReLU) apply pass."""
    w = P[wname]
    ws = K.fused_stats_workspace(x, w) if training else None
    r = K.conv1x1_nhwc(x, w, stride, stats=ws, sink_w=_sink(G, wname), passthrough=passthrough)
    o, xp = r if passthrough else (r, None)
    y = _bn(o, P, A, G, bnname, relu, residual, training, eps, momentum, pre_stats=ws)
    return (y, xp) if passthrough else y


# ----------------------------------------------------------------------------- inference: BatchNorm folded away
_FOLD_CACHE = {}


def _folded(P, A, wname, bnname, eps):
    """(w', shift): convolution weights with the BatchNorm scale folded in (w' = w * gamma / sqrt(var + eps) per output
    channel) and the remaining per-channel shift, cached until any of the five tensors is modified in place."""
    w, g, b = P[wname], P[bnname + "_gamma"], P[bnname + "_beta"]
    m, v = A[bnname + "_running_mean"], A[bnname + "_running_var"]
    key = (w.data_ptr(), g.data_ptr(), b.data_ptr(), m.data_ptr(), v.data_ptr())
    ver = (w._version, g._version, b._version, m._version, v._version, float(eps))
    hit = _FOLD_CACHE.get(key)
    # the entry must belong to THESE tensor objects: after a model is freed the caching allocator hands the same addresses
    # (with the same fresh version counters) to the next model's parameters
    if hit is not None and hit[0] == ver and all(r() is t for r, t in zip(hit[3], (w, g, b, m, v))):
        return hit[1], hit[2]
    with torch.no_grad():
        scale = g.float() * torch.rsqrt(v.float() + eps)
        shift = (b.float() - m.float() * scale).contiguous()
        wf = (w.float() * scale.view(-1, *([1] * (w.dim() - 1)))).to(w.dtype)
        if w.dim() == 4:
            wf = wf.contiguous(memory_format=torch.channels_last)
        else:
            wf = wf.contiguous()
    if len(_FOLD_CACHE) > 4096:
        _FOLD_CACHE.clear()
    _FOLD_CACHE[key] = (ver, wf, shift, tuple(weakref.ref(t) for t in (w, g, b, m, v)))
    return wf, shift


def _conv_bn_eval(x, P, A, wname, bnname, relu, stride=1, pad=0, residual=None, eps=1e-5):
    """Inference Conv+BN(+residual)(+ReLU) as ONE kernel: the BN scale lives in the weights, the shift is the GEMM
    bias, residual add and ReLU run in the GEMM epilogue (1x1) / in place after the library conv (kxk)."""
    wf, shift = _folded(P, A, wname, bnname, eps)
    if wf.dim() == 2:                                        # 1x1: tcgen05 GEMM
        xs = K.subsample_nhwc(x, stride)
        n, h, ww, c = xs.shape
        y = K.gemm_bias_act(xs.reshape(n * h * ww, c), wf, shift, relu=relu,
                            addend=residual.reshape(n * h * ww, -1) if residual is not None else None)
        return y.view(n, h, ww, wf.shape[0])
    if K.stem_conv_ok(x, wf):
        y = K.stem_conv_nhwc(x, wf, stride, pad, bias=shift, relu=relu)
        return y
    if K.conv_nhwc_ok(x, wf):
        # implicit-GEMM tcgen05 convolution: folded-BN shift = bias, residual add and ReLU in the epilogue
        return K.conv_fwd_raw(x.contiguous(), wf.permute(0, 2, 3, 1), stride, pad, bias=shift, relu=relu,
                              addend=residual.contiguous() if residual is not None else None)
    y = F.conv2d(x.permute(0, 3, 1, 2), wf, shift.to(wf.dtype), stride, pad).permute(0, 2, 3, 1)
    if residual is not None:
        y = y + residual
    return torch.relu_(y) if relu else y


def _resnet_forward_eval(spec, P, A, x, eps):
    st, pd = (1, 1) if spec.thumbnail else (2, 3)
    y = _conv_bn_eval(x, P, A, "conv0_weight", "bn0", True, st, pd, eps=eps)
    if not spec.thumbnail:
        y = K.max_pool_nhwc(y.contiguous(), 3, 2, 1)
    for s, (n, cout) in enumerate(zip(spec.layers, spec.channels[1:])):
        for b in range(n):
            pre = "stage%d_unit%d_" % (s + 1, b + 1)
            stride = 2 if (b == 0 and s > 0) else 1
            has_down = (pre + "down_weight") in P
            res = _conv_bn_eval(y, P, A, pre + "down_weight", pre + "down_bn", False, stride, eps=eps) if has_down else y
            if spec.block == "bottleneck":
                s1, s2 = (stride, 1) if spec.version == "v1" else (1, stride)
                o = _conv_bn_eval(y, P, A, pre + "conv1_weight", pre + "bn1", True, s1, eps=eps)
                o = _conv_bn_eval(o, P, A, pre + "conv2_weight", pre + "bn2", True, s2, 1, eps=eps)
                y = _conv_bn_eval(o.contiguous(), P, A, pre + "conv3_weight", pre + "bn3", True, 1, residual=res.contiguous(), eps=eps)
            else:
                o = _conv_bn_eval(y, P, A, pre + "conv1_weight", pre + "bn1", True, stride, 1, eps=eps)
                y = _conv_bn_eval(o, P, A, pre + "conv2_weight", pre + "bn2", True, 1, 1, residual=res, eps=eps)
    pooled = y.float().mean(dim=(1, 2)).to(y.dtype)
    return K.linear(pooled, P["fc_weight"], P["fc_bias"])


def resnet_forward(spec, P, A, x, training=True, eps=1e-5, momentum=0.9, G=None, stem_pad=None):
    """x: [N, H, W, C] channels-last.  Returns logits [N, classes] (bf16).

    G (optional): name -> fp32 gradient-arena view; gradients of 1x1-conv / FC weights and BN gamma/beta are
    then written straight into the arena by the backward kernels."""
    if (not training and not torch.is_grad_enabled() and x.is_cuda and x.dtype is torch.bfloat16
            and getenv("MXB200_FOLD_BN", True)):
        return _resnet_forward_eval(spec, P, A, x, eps)
    if stem_pad is None:
        stem_pad = int(getenv("MXB200_STEM_PAD", 0))   # measured on B200: padding 3->8 channels is slower (2.7 vs 1.4 ms fwd)
    kw = dict(training=training, eps=eps, momentum=momentum)
    st, pd = (1, 1) if spec.thumbnail else (2, 3)
    w0 = P["conv0_weight"]
    if K.stem_conv_ok(x, w0):
        # stem as im2col + tcgen05 GEMM: BN statistics come out of the GEMM epilogue, dW goes straight to the arena
        ws = K.fused_stats_workspace(x.reshape(-1, x.shape[-1]), w0.reshape(w0.shape[0], -1), force=True) if training else None
        y = K.stem_conv_nhwc(x, w0, st, pd, stats=ws, sink_w=_sink(G, "conv0_weight"))
        y = _bn(y, P, A, G, "bn0", True, pre_stats=ws, **kw)
    else:
        y = _stem_conv(x, w0, st, pd, stem_pad)
        y = _bn(y.contiguous(), P, A, G, "bn0", True, **kw)
    if not spec.thumbnail:
        y = K.max_pool_nhwc(y, 3, 2, 1)
    for s, (n, cout) in enumerate(zip(spec.layers, spec.channels[1:])):
        for b in range(n):
            pre = "stage%d_unit%d_" % (s + 1, b + 1)
            stride = 2 if (b == 0 and s > 0) else 1
            has_down = (pre + "down_weight") in P
            if spec.block == "bottleneck":
                # v1b: stride lives on the 3x3 (v1 puts it on the first 1x1)
                s1, s2 = (stride, 1) if spec.version == "v1" else (1, stride)
                if s1 == 1:
                    # conv1 also returns its input so the skip branch hangs off it: the skip gradient is then
                    # added inside conv1's dgrad GEMM epilogue instead of a separate elementwise add
                    o, skip = _conv1x1_bn(y, P, A, G, pre + "conv1_weight", pre + "bn1", True, 1, passthrough=True, **kw)
                else:
                    o = _conv1x1_bn(y, P, A, G, pre + "conv1_weight", pre + "bn1", True, s1, **kw)
                    skip = y
                res = _conv1x1_bn(skip, P, A, G, pre + "down_weight", pre + "down_bn", False, stride, **kw) if has_down else skip
                o = _conv_bn(o, P, A, G, pre + "conv2_weight", pre + "bn2", True, s2, 1, **kw)
                y = _conv1x1_bn(o, P, A, G, pre + "conv3_weight", pre + "bn3", True, 1, residual=res, **kw)
            else:
                res = _conv1x1_bn(y, P, A, G, pre + "down_weight", pre + "down_bn", False, stride, **kw) if has_down else y
                o = _conv_bn(y, P, A, G, pre + "conv1_weight", pre + "bn1", True, stride, 1, **kw)
                y = _conv_bn(o, P, A, G, pre + "conv2_weight", pre + "bn2", True, 1, 1, residual=res, **kw)
    pooled = y.float().mean(dim=(1, 2)).to(y.dtype)  # global average pool
    return K.linear(pooled, P["fc_weight"], P["fc_bias"], sink_w=_sink(G, "fc_weight"), sink_b=_sink(G, "fc_bias"))
