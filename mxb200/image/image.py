"""imdecode / imread / resize / crops / augmenters / ImageIter (reference: python/mxnet/image/image.py).

Provenance: adapted from Apache MXNet (incubating) 1.3.1 ``python/mxnet/image/image.py`` (Apache License 2.0).  The public API, argument
names and control flow follow the original so that user scripts keep working; the tensor maths runs on torch /
the mxb200 kernels.  This file is API-compatibility glue, not original B200 work ("Provenance").
"""
import os
import random

import numpy as np
import torch

from ..ndarray import NDArray, array
from .. import io as _io
from .. import recordio

__all__ = ["imdecode", "imread", "imresize", "resize_short", "fixed_crop", "random_crop", "center_crop", "color_normalize",
           "scale_down", "Augmenter", "ResizeAug", "ForceResizeAug", "RandomCropAug", "CenterCropAug", "HorizontalFlipAug",
           "CastAug", "ColorNormalizeAug", "CreateAugmenter", "ImageIter", "random_size_crop", "SequentialAug", "RandomSizedCropAug", "RandomOrderAug", "BrightnessJitterAug", "ContrastJitterAug", "SaturationJitterAug", "HueJitterAug", "ColorJitterAug", "LightingAug", "RandomGrayAug"]


def imdecode(buf, flag=1, to_rgb=1, out=None):
    import cv2
    if isinstance(buf, NDArray):
        buf = buf.asnumpy().tobytes()
    img = cv2.imdecode(np.frombuffer(buf, dtype=np.uint8), cv2.IMREAD_COLOR if flag == 1 else cv2.IMREAD_GRAYSCALE)
    if img is None:
        raise ValueError("imdecode: could not decode image buffer")
    if flag == 1 and to_rgb:
        img = img[:, :, ::-1]
    if img.ndim == 2:
        img = img[:, :, None]
    return array(np.ascontiguousarray(img), dtype=np.uint8)


def imread(filename, flag=1, to_rgb=1):
    if filename.endswith(".npy"):
        return array(np.load(filename), dtype=np.uint8)
    import os
    if not os.path.isfile(filename):
        from ..base import MXNetError
        raise MXNetError("imread: cannot open %s" % filename)      # same error class as the C++ reader
    with open(filename, "rb") as f:
        return imdecode(f.read(), flag, to_rgb)


def imresize(src, w, h, interp=1):
    import cv2
    a = src.asnumpy()
    r = cv2.resize(a, (w, h), interpolation={0: cv2.INTER_NEAREST, 1: cv2.INTER_LINEAR, 2: cv2.INTER_CUBIC,
                                             3: cv2.INTER_AREA, 4: cv2.INTER_LANCZOS4}.get(interp, cv2.INTER_LINEAR))
    if r.ndim == 2:
        r = r[:, :, None]
    return array(r, dtype=a.dtype)


def scale_down(src_size, size):
    w, h = size
    sw, sh = src_size
    if sh < h:
        w, h = float(w * sh) / h, sh
    if sw < w:
        w, h = sw, float(h * sw) / w
    return int(w), int(h)


def resize_short(src, size, interp=2):
    h, w, _ = src.shape
    if h > w:
        new_h, new_w = size * h // w, size
    else:
        new_h, new_w = size, size * w // h
    return imresize(src, new_w, new_h, interp=interp)


def fixed_crop(src, x0, y0, w, h, size=None, interp=2):
    out = src[y0:y0 + h, x0:x0 + w]
    if size is not None and (w, h) != size:
        out = imresize(out, *size, interp=interp)
    return out


def random_crop(src, size, interp=2):
    h, w, _ = src.shape
    new_w, new_h = scale_down((w, h), size)
    x0 = random.randint(0, w - new_w)
    y0 = random.randint(0, h - new_h)
    return fixed_crop(src, x0, y0, new_w, new_h, size, interp), (x0, y0, new_w, new_h)


def center_crop(src, size, interp=2):
    h, w, _ = src.shape
    new_w, new_h = scale_down((w, h), size)
    x0 = int((w - new_w) / 2)
    y0 = int((h - new_h) / 2)
    return fixed_crop(src, x0, y0, new_w, new_h, size, interp), (x0, y0, new_w, new_h)


def color_normalize(src, mean, std=None):
    if mean is not None:
        src = src - mean
    if std is not None:
        src = src / std
    return src


class Augmenter(object):
    def __init__(self, **kwargs):
        self._kwargs = kwargs

    def dumps(self):
        import json

        def conv(v):
            if isinstance(v, np.ndarray):
                return v.tolist()
            if isinstance(v, NDArray):
                return v.asnumpy().tolist()
            return v
        return json.dumps([self.__class__.__name__.lower(), {k: conv(v) for k, v in self._kwargs.items()}])

    def __call__(self, src):
        raise NotImplementedError("Must override implementation.")


class ResizeAug(Augmenter):
    def __init__(self, size, interp=2):
        super(ResizeAug, self).__init__(size=size, interp=interp)
        self.size, self.interp = size, interp

    def __call__(self, src):
        return resize_short(src, self.size, self.interp)


class ForceResizeAug(Augmenter):
    def __init__(self, size, interp=2):
        super(ForceResizeAug, self).__init__(size=size, interp=interp)
        self.size, self.interp = size, interp

    def __call__(self, src):
        return imresize(src, *self.size, interp=self.interp)


class RandomCropAug(Augmenter):
    def __init__(self, size, interp=2):
        super(RandomCropAug, self).__init__(size=size, interp=interp)
        self.size, self.interp = size, interp

    def __call__(self, src):
        return random_crop(src, self.size, self.interp)[0]


class CenterCropAug(Augmenter):
    def __init__(self, size, interp=2):
        super(CenterCropAug, self).__init__(size=size, interp=interp)
        self.size, self.interp = size, interp

    def __call__(self, src):
        return center_crop(src, self.size, self.interp)[0]


class HorizontalFlipAug(Augmenter):
    def __init__(self, p):
        super(HorizontalFlipAug, self).__init__(p=p)
        self.p = p

    def __call__(self, src):
        if random.random() < self.p:
            return NDArray(torch.flip(src._t, [1]), src.context)
        return src


class CastAug(Augmenter):
    def __init__(self, typ="float32"):
        super(CastAug, self).__init__(type=typ)
        self.typ = typ

    def __call__(self, src):
        return src.astype(self.typ)


class ColorNormalizeAug(Augmenter):
    def __init__(self, mean, std):
        super(ColorNormalizeAug, self).__init__(mean=mean, std=std)
        self.mean = array(mean) if mean is not None else None
        self.std = array(std) if std is not None else None

    def __call__(self, src):
        return color_normalize(src, self.mean, self.std)


def random_size_crop(src, size, area, ratio, interp=2, **kwargs):
    """Random crop with area in ``area`` (fraction or (min, max)) and aspect ratio in ``ratio``, resized to ``size``
    (Inception-style; image.py:441-495).  Falls back to a centre crop after 10 failed draws."""
    h, w = src.shape[0], src.shape[1]
    src_area = h * w
    if "min_area" in kwargs:
        area = kwargs.pop("min_area")
    if not isinstance(area, (tuple, list)):
        area = (area, 1.0)
    for _ in range(10):
        target_area = random.uniform(area[0], area[1]) * src_area
        log_ratio = (np.log(ratio[0]), np.log(ratio[1]))
        new_ratio = np.exp(random.uniform(*log_ratio))
        new_w = int(round(np.sqrt(target_area * new_ratio)))
        new_h = int(round(np.sqrt(target_area / new_ratio)))
        if new_w <= w and new_h <= h:
            x0 = random.randint(0, w - new_w)
            y0 = random.randint(0, h - new_h)
            return fixed_crop(src, x0, y0, new_w, new_h, size, interp), (x0, y0, new_w, new_h)
    return center_crop(src, size, interp)


class SequentialAug(Augmenter):
    def __init__(self, ts):
        super(SequentialAug, self).__init__()
        self.ts = ts

    def dumps(self):
        return [self.__class__.__name__.lower(), [x.dumps() for x in self.ts]]

    def __call__(self, src):
        for aug in self.ts:
            src = aug(src)
        return src


class RandomSizedCropAug(Augmenter):
    def __init__(self, size, area, ratio, interp=2, **kwargs):
        super(RandomSizedCropAug, self).__init__(size=size, area=area, ratio=ratio, interp=interp)
        self.size, self.area, self.ratio, self.interp = size, kwargs.pop("min_area", area), ratio, interp

    def __call__(self, src):
        return random_size_crop(src, self.size, self.area, self.ratio, self.interp)[0]


class RandomOrderAug(Augmenter):
    def __init__(self, ts):
        super(RandomOrderAug, self).__init__()
        self.ts = ts

    def dumps(self):
        return [self.__class__.__name__.lower(), [x.dumps() for x in self.ts]]

    def __call__(self, src):
        random.shuffle(self.ts)
        for t in self.ts:
            src = t(src)
        return src


_GRAY = np.array([[[0.299, 0.587, 0.114]]], dtype=np.float32)


class BrightnessJitterAug(Augmenter):
    def __init__(self, brightness):
        super(BrightnessJitterAug, self).__init__(brightness=brightness)
        self.brightness = brightness

    def __call__(self, src):
        alpha = 1.0 + random.uniform(-self.brightness, self.brightness)
        return src * alpha


class ContrastJitterAug(Augmenter):
    def __init__(self, contrast):
        super(ContrastJitterAug, self).__init__(contrast=contrast)
        self.contrast = contrast

    def __call__(self, src):
        alpha = 1.0 + random.uniform(-self.contrast, self.contrast)
        gray = (src.astype("float32") * array(_GRAY)).sum()
        gray = (3.0 * (1.0 - alpha) / float(src.size)) * float(gray.asnumpy().reshape(-1)[0])
        return src * alpha + gray


class SaturationJitterAug(Augmenter):
    def __init__(self, saturation):
        super(SaturationJitterAug, self).__init__(saturation=saturation)
        self.saturation = saturation

    def __call__(self, src):
        alpha = 1.0 + random.uniform(-self.saturation, self.saturation)
        gray = (src.astype("float32") * array(_GRAY)).sum(axis=2, keepdims=True) * (1.0 - alpha)
        return src * alpha + gray


class HueJitterAug(Augmenter):
    """Rotate the hue in YIQ space (image.py:752-783)."""

    def __init__(self, hue):
        super(HueJitterAug, self).__init__(hue=hue)
        self.hue = hue
        self.tyiq = np.array([[0.299, 0.587, 0.114], [0.596, -0.274, -0.321], [0.211, -0.523, 0.311]])
        self.ityiq = np.array([[1.0, 0.956, 0.621], [1.0, -0.272, -0.647], [1.0, -1.107, 1.705]])

    def __call__(self, src):
        alpha = random.uniform(-self.hue, self.hue)
        u, w = np.cos(alpha * np.pi), np.sin(alpha * np.pi)
        bt = np.array([[1.0, 0.0, 0.0], [0.0, u, -w], [0.0, w, u]])
        t = np.dot(np.dot(self.ityiq, bt), self.tyiq).T
        return NDArray((src._t.float() @ torch.from_numpy(t).float().to(src._t.device)), src.context)


class ColorJitterAug(RandomOrderAug):
    def __init__(self, brightness, contrast, saturation):
        ts = []
        if brightness > 0:
            ts.append(BrightnessJitterAug(brightness))
        if contrast > 0:
            ts.append(ContrastJitterAug(contrast))
        if saturation > 0:
            ts.append(SaturationJitterAug(saturation))
        super(ColorJitterAug, self).__init__(ts)


class LightingAug(Augmenter):
    """AlexNet-style PCA noise (image.py:809-832)."""

    def __init__(self, alphastd, eigval, eigvec):
        super(LightingAug, self).__init__(alphastd=alphastd, eigval=eigval, eigvec=eigvec)
        self.alphastd, self.eigval, self.eigvec = alphastd, np.asarray(eigval), np.asarray(eigvec)

    def __call__(self, src):
        alpha = np.random.normal(0, self.alphastd, size=(3,))
        rgb = np.dot(self.eigvec * alpha, self.eigval)
        return src + array(rgb.astype(np.float32))


class RandomGrayAug(Augmenter):
    def __init__(self, p):
        super(RandomGrayAug, self).__init__(p=p)
        self.p = p
        self.mat = torch.tensor([[0.21, 0.21, 0.21], [0.72, 0.72, 0.72], [0.07, 0.07, 0.07]])

    def __call__(self, src):
        if random.random() < self.p:
            return NDArray(src._t.float() @ self.mat.to(src._t.device), src.context)
        return src


def CreateAugmenter(data_shape, resize=0, rand_crop=False, rand_resize=False, rand_mirror=False, mean=None, std=None,
                    brightness=0, contrast=0, saturation=0, hue=0, pca_noise=0, rand_gray=0, inter_method=2):
    """Standard augmenter list (image.py:908-1019): resize, (random sized | random | centre) crop, mirror, cast, colour
    jitter, hue, PCA lighting, random gray, normalisation."""
    auglist = []
    if resize > 0:
        auglist.append(ResizeAug(resize, inter_method))
    crop_size = (data_shape[2], data_shape[1])
    if rand_resize:
        assert rand_crop
        auglist.append(RandomSizedCropAug(crop_size, 0.08, (3.0 / 4.0, 4.0 / 3.0), inter_method))
    elif rand_crop:
        auglist.append(RandomCropAug(crop_size, inter_method))
    else:
        auglist.append(CenterCropAug(crop_size, inter_method))
    if rand_mirror:
        auglist.append(HorizontalFlipAug(0.5))
    auglist.append(CastAug())
    if brightness or contrast or saturation:
        auglist.append(ColorJitterAug(brightness, contrast, saturation))
    if hue:
        auglist.append(HueJitterAug(hue))
    if pca_noise > 0:
        eigval = np.array([55.46, 4.794, 1.148])
        eigvec = np.array([[-0.5675, 0.7192, 0.4009], [-0.5808, -0.0045, -0.8140], [-0.5836, -0.6948, 0.4203]])
        auglist.append(LightingAug(pca_noise, eigval, eigvec))
    if rand_gray > 0:
        auglist.append(RandomGrayAug(rand_gray))
    if mean is True:
        mean = np.array([123.68, 116.28, 103.53])
    if std is True:
        std = np.array([58.395, 57.12, 57.375])
    if mean is not None or std is not None:
        auglist.append(ColorNormalizeAug(mean, std))
    return auglist


class ImageIter(_io.DataIter):
    """Image iterator over a RecordIO pack or an image list, with python augmenters (image.py:ImageIter)."""

    def __init__(self, batch_size, data_shape, label_width=1, path_imgrec=None, path_imglist=None, path_root=None,
                 path_imgidx=None, shuffle=False, part_index=0, num_parts=1, aug_list=None, imglist=None,
                 data_name="data", label_name="softmax_label", dtype="float32", last_batch_handle="pad", **kwargs):
        super(ImageIter, self).__init__(batch_size)
        assert path_imgrec or path_imglist or isinstance(imglist, list)
        self.imgrec = None
        self.seq = None
        self.imglist = None
        if path_imgrec:
            if path_imgidx:
                self.imgrec = recordio.MXIndexedRecordIO(path_imgidx, path_imgrec, "r")
                self.seq = list(self.imgrec.keys)
            else:
                self.imgrec = recordio.MXRecordIO(path_imgrec, "r")
        if path_imglist:
            self.imglist = {}
            self.seq = []
            with open(path_imglist) as fin:
                for line in fin:
                    line = line.strip().split("\t")
                    label = np.array(line[1:-1], dtype=np.float32)
                    self.imglist[int(line[0])] = (label, line[-1])
                    self.seq.append(int(line[0]))
        elif isinstance(imglist, list):
            self.imglist = {}
            self.seq = []
            for i, img in enumerate(imglist):
                self.imglist[i] = (np.array(img[0] if isinstance(img[0], (list, np.ndarray)) else [img[0]], dtype=np.float32), img[1])
                self.seq.append(i)
        self.path_root = path_root
        self.data_shape = tuple(data_shape)
        self.label_width = label_width
        self.shuffle = shuffle
        if self.seq is not None and num_parts > 1:
            n = len(self.seq) // num_parts
            self.seq = self.seq[part_index * n:(part_index + 1) * n]
        self.auglist = aug_list if aug_list is not None else CreateAugmenter(data_shape, **kwargs)
        self.provide_data = [_io.DataDesc(data_name, (batch_size,) + self.data_shape, dtype)]
        self.provide_label = [_io.DataDesc(label_name, (batch_size, label_width) if label_width > 1 else (batch_size,), dtype)]
        self.cur = 0
        self.reset()

    def reset(self):
        if self.shuffle and self.seq is not None:
            random.shuffle(self.seq)
        if self.imgrec is not None:
            self.imgrec.reset()
        self.cur = 0

    def next_sample(self):
        if self.seq is not None:
            if self.cur >= len(self.seq):
                raise StopIteration
            idx = self.seq[self.cur]
            self.cur += 1
            if self.imgrec is not None:
                s = self.imgrec.read_idx(idx)
                header, img = recordio.unpack(s)
                return (header.label if self.imglist is None else self.imglist[idx][0]), imdecode(img)
            label, fname = self.imglist[idx]
            return label, imread(os.path.join(self.path_root or "", fname))
        s = self.imgrec.read()
        if s is None:
            raise StopIteration
        header, img = recordio.unpack(s)
        return header.label, imdecode(img)

    def next(self):
        bs = self.batch_size
        c, h, w = self.data_shape
        data = np.zeros((bs, c, h, w), dtype=np.float32)
        label = np.zeros((bs, self.label_width) if self.label_width > 1 else (bs,), dtype=np.float32)
        i = 0
        try:
            while i < bs:
                lab, img = self.next_sample()
                for aug in self.auglist:
                    img = aug(img)
                data[i] = img.asnumpy().transpose(2, 0, 1)
                label[i] = lab
                i += 1
        except StopIteration:
            if i == 0:
                raise
        return _io.DataBatch([array(data)], [array(label)], pad=bs - i)
