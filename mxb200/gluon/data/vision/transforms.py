"""Image transforms as Blocks (reference: python/mxnet/gluon/data/vision/transforms.py)."""
import numpy as np
import torch

from ...block import Block, HybridBlock
from ...nn import Sequential, HybridSequential
from ....ndarray import NDArray
from ....ndarray.ndarray import _op

__all__ = ["Compose", "Cast", "ToTensor", "Normalize", "RandomResizedCrop", "CenterCrop", "Resize", "RandomFlipLeftRight",
           "RandomFlipTopBottom", "RandomBrightness", "RandomContrast", "RandomSaturation", "RandomHue", "RandomColorJitter",
           "RandomLighting"]


class Compose(Sequential):
    def __init__(self, transforms):
        super(Compose, self).__init__()
        for t in transforms:
            self.add(t)


class Cast(HybridBlock):
    def __init__(self, dtype="float32"):
        super(Cast, self).__init__()
        self._dtype = dtype

    def hybrid_forward(self, F, x):
        return F.cast(x, self._dtype)


class ToTensor(HybridBlock):
    """(H,W,C) uint8 [0,255] -> (C,H,W) float32 [0,1)."""

    def hybrid_forward(self, F, x):
        return _op(lambda t: (t.float() / 255.0).permute(2, 0, 1) if t.dim() == 3 else (t.float() / 255.0).permute(0, 3, 1, 2), x)


class Normalize(HybridBlock):
    def __init__(self, mean, std):
        super(Normalize, self).__init__()
        self._mean, self._std = mean, std

    def hybrid_forward(self, F, x):
        def f(t):
            m = torch.as_tensor(self._mean, dtype=t.dtype, device=t.device).reshape(-1, 1, 1)
            s = torch.as_tensor(self._std, dtype=t.dtype, device=t.device).reshape(-1, 1, 1)
            return (t - m) / s
        return _op(f, x)


def _resize(t, size):
    import torch.nn.functional as F
    hwc = t.dim() == 3
    x = t.permute(2, 0, 1)[None].float() if hwc else t.float()
    y = F.interpolate(x, size=size, mode="bilinear", align_corners=False)
    return y[0].permute(1, 2, 0).to(t.dtype) if hwc else y.to(t.dtype)


class Resize(Block):
    def __init__(self, size, keep_ratio=False, interpolation=1):
        super(Resize, self).__init__()
        self._size = (size, size) if isinstance(size, int) else tuple(size)

    def forward(self, x):
        return _op(lambda t: _resize(t, (self._size[1], self._size[0])), x)


class CenterCrop(Block):
    def __init__(self, size, interpolation=1):
        super(CenterCrop, self).__init__()
        self._size = (size, size) if isinstance(size, int) else tuple(size)

    def forward(self, x):
        def f(t):
            h, w = t.shape[0], t.shape[1]
            cw, ch = self._size
            y0, x0 = max(0, (h - ch) // 2), max(0, (w - cw) // 2)
            return t[y0:y0 + ch, x0:x0 + cw]
        return _op(f, x)


class RandomResizedCrop(Block):
    def __init__(self, size, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0), interpolation=1):
        super(RandomResizedCrop, self).__init__()
        self._size = (size, size) if isinstance(size, int) else tuple(size)
        self._scale, self._ratio = scale, ratio

    def forward(self, x):
        def f(t):
            h, w = t.shape[0], t.shape[1]
            area = h * w
            for _ in range(10):
                ta = np.random.uniform(*self._scale) * area
                ar = np.exp(np.random.uniform(np.log(self._ratio[0]), np.log(self._ratio[1])))
                cw, ch = int(round(np.sqrt(ta * ar))), int(round(np.sqrt(ta / ar)))
                if cw <= w and ch <= h:
                    x0, y0 = np.random.randint(0, w - cw + 1), np.random.randint(0, h - ch + 1)
                    return _resize(t[y0:y0 + ch, x0:x0 + cw], (self._size[1], self._size[0]))
            return _resize(t, (self._size[1], self._size[0]))
        return _op(f, x)


class RandomFlipLeftRight(HybridBlock):
    def hybrid_forward(self, F, x):
        return _op(lambda t: torch.flip(t, [1]) if np.random.rand() < 0.5 else t, x)


class RandomFlipTopBottom(HybridBlock):
    def hybrid_forward(self, F, x):
        return _op(lambda t: torch.flip(t, [0]) if np.random.rand() < 0.5 else t, x)


class _ImageOp(HybridBlock):
    _op_name, _arg_names = None, ()

    def __init__(self, *args):
        super(_ImageOp, self).__init__()
        self._args = dict(zip(self._arg_names, args))

    def hybrid_forward(self, F, x):
        return getattr(F.image, self._op_name)(x, **self._args)


def _jitter_range(v):
    return max(0, 1 - v), 1 + v


class RandomBrightness(_ImageOp):
    """Scale brightness by a factor drawn from [max(0, 1 - b), 1 + b] (transforms.py:RandomBrightness of the reference)."""
    _op_name, _arg_names = "random_brightness", ("min_factor", "max_factor")

    def __init__(self, brightness):
        super(RandomBrightness, self).__init__(*_jitter_range(brightness))


class RandomContrast(_ImageOp):
    _op_name, _arg_names = "random_contrast", ("min_factor", "max_factor")

    def __init__(self, contrast):
        super(RandomContrast, self).__init__(*_jitter_range(contrast))


class RandomSaturation(_ImageOp):
    _op_name, _arg_names = "random_saturation", ("min_factor", "max_factor")

    def __init__(self, saturation):
        super(RandomSaturation, self).__init__(*_jitter_range(saturation))


class RandomHue(_ImageOp):
    _op_name, _arg_names = "random_hue", ("min_factor", "max_factor")

    def __init__(self, hue):
        super(RandomHue, self).__init__(*_jitter_range(hue))


class RandomColorJitter(_ImageOp):
    _op_name, _arg_names = "random_color_jitter", ("brightness", "contrast", "saturation", "hue")

    def __init__(self, brightness=0, contrast=0, saturation=0, hue=0):
        super(RandomColorJitter, self).__init__(brightness, contrast, saturation, hue)


class RandomLighting(_ImageOp):
    _op_name, _arg_names = "random_lighting", ("alpha_std",)

    def __init__(self, alpha):
        super(RandomLighting, self).__init__(alpha)
