"""Detection augmenters and iterator (reference: python/mxnet/image/detection.py).  Labels are [N, 5+] arrays of
(class, xmin, ymin, xmax, ymax, ...) with coordinates normalised to [0, 1]; -1 rows are padding."""
import json
import random

import numpy as np

from .. import io as _io
from ..ndarray import array
from . import image as _img
from .image import Augmenter, ImageIter

__all__ = ["DetAugmenter", "DetBorrowAug", "DetRandomSelectAug", "DetHorizontalFlipAug", "DetRandomCropAug",
           "DetRandomPadAug", "CreateMultiRandCropAugmenter", "CreateDetAugmenter", "ImageDetIter"]


class DetAugmenter(object):
    def __init__(self, **kwargs):
        self._kwargs = kwargs

    def dumps(self):
        return json.dumps([self.__class__.__name__.lower(), self._kwargs])

    def __call__(self, src, label):
        raise NotImplementedError("Must override implementation.")


class DetBorrowAug(DetAugmenter):
    """Apply an image-only augmenter, labels untouched."""

    def __init__(self, augmenter):
        if not isinstance(augmenter, Augmenter):
            raise TypeError("Borrowing from invalid Augmenter")
        super(DetBorrowAug, self).__init__(augmenter=augmenter.dumps())
        self.augmenter = augmenter

    def __call__(self, src, label):
        return self.augmenter(src), label


class DetRandomSelectAug(DetAugmenter):
    """Pick one augmenter of the list at random (or none with probability ``skip_prob``)."""

    def __init__(self, aug_list, skip_prob=0):
        super(DetRandomSelectAug, self).__init__(skip_prob=skip_prob)
        self.aug_list = aug_list if isinstance(aug_list, (list, tuple)) else [aug_list]
        self.skip_prob = skip_prob if self.aug_list else 1

    def __call__(self, src, label):
        if random.random() < self.skip_prob:
            return src, label
        return random.choice(self.aug_list)(src, label)


class DetHorizontalFlipAug(DetAugmenter):
    def __init__(self, p):
        super(DetHorizontalFlipAug, self).__init__(p=p)
        self.p = p

    def __call__(self, src, label):
        if random.random() < self.p:
            src = _img.HorizontalFlipAug(1.0)(src)
            label = label.copy()
            valid = label[:, 0] >= 0
            xmin = 1.0 - label[valid, 3]
            label[valid, 3] = 1.0 - label[valid, 1]
            label[valid, 1] = xmin
        return src, label


def _box_area(b):
    return np.maximum(0, b[:, 2] - b[:, 0]) * np.maximum(0, b[:, 3] - b[:, 1])


class DetRandomCropAug(DetAugmenter):
    """Random crop constrained by object coverage (detection.py:152-320): keeps boxes whose centre lies in the crop and
    whose visible fraction is at least ``min_eject_coverage``."""

    def __init__(self, min_object_covered=0.1, aspect_ratio_range=(0.75, 1.33), area_range=(0.05, 1.0),
                 min_eject_coverage=0.3, max_attempts=50):
        super(DetRandomCropAug, self).__init__(min_object_covered=min_object_covered, aspect_ratio_range=aspect_ratio_range,
                                               area_range=area_range, min_eject_coverage=min_eject_coverage,
                                               max_attempts=max_attempts)
        self.min_object_covered, self.aspect_ratio_range = min_object_covered, aspect_ratio_range
        self.area_range, self.min_eject_coverage, self.max_attempts = area_range, min_eject_coverage, max_attempts
        self.enabled = area_range[1] > 0 and area_range[0] <= area_range[1] and aspect_ratio_range[0] <= aspect_ratio_range[1]

    def _propose(self, label, h, w):
        valid = label[label[:, 0] >= 0]
        for _ in range(self.max_attempts):
            ratio = random.uniform(*self.aspect_ratio_range)
            area = random.uniform(*self.area_range)
            cw = min(1.0, np.sqrt(area * ratio))
            ch = min(1.0, np.sqrt(area / ratio))
            x0, y0 = random.uniform(0, 1 - cw), random.uniform(0, 1 - ch)
            crop = np.array([x0, y0, x0 + cw, y0 + ch])
            if int(cw * w) < 1 or int(ch * h) < 1:
                continue
            if valid.shape[0] == 0:
                return crop, label
            b = valid[:, 1:5]
            inter = np.concatenate([np.maximum(b[:, :2], crop[:2]), np.minimum(b[:, 2:4], crop[2:])], axis=1)
            cover = _box_area(inter) / np.maximum(_box_area(b), 1e-12)
            if cover.max() < self.min_object_covered:
                continue
            cx, cy = (b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2
            keep = (cx > crop[0]) & (cx < crop[2]) & (cy > crop[1]) & (cy < crop[3]) & (cover >= self.min_eject_coverage)
            if not keep.any():
                continue
            new = valid[keep].copy()
            nb = inter[keep]
            new[:, 1] = (nb[:, 0] - x0) / cw
            new[:, 2] = (nb[:, 1] - y0) / ch
            new[:, 3] = (nb[:, 2] - x0) / cw
            new[:, 4] = (nb[:, 3] - y0) / ch
            out = np.full_like(label, -1.0)
            out[:new.shape[0]] = new
            return crop, out
        return None, label

    def __call__(self, src, label):
        if not self.enabled:
            return src, label
        h, w = src.shape[0], src.shape[1]
        crop, new_label = self._propose(label, h, w)
        if crop is None:
            return src, label
        x0, y0 = int(crop[0] * w), int(crop[1] * h)
        cw, ch = max(1, int((crop[2] - crop[0]) * w)), max(1, int((crop[3] - crop[1]) * h))
        return _img.fixed_crop(src, x0, y0, cw, ch), new_label


class DetRandomPadAug(DetAugmenter):
    """Place the image on a larger canvas filled with ``pad_val`` (zoom-out; detection.py:323-414)."""

    def __init__(self, aspect_ratio_range=(0.75, 1.33), area_range=(1.0, 3.0), max_attempts=50, pad_val=(128, 128, 128)):
        super(DetRandomPadAug, self).__init__(aspect_ratio_range=aspect_ratio_range, area_range=area_range,
                                              max_attempts=max_attempts, pad_val=pad_val)
        self.aspect_ratio_range, self.area_range, self.max_attempts = aspect_ratio_range, area_range, max_attempts
        self.pad_val = pad_val if isinstance(pad_val, (list, tuple)) else (pad_val,) * 3
        self.enabled = area_range[1] > 1.0

    def __call__(self, src, label):
        if not self.enabled:
            return src, label
        h, w = src.shape[0], src.shape[1]
        for _ in range(self.max_attempts):
            ratio = random.uniform(*self.aspect_ratio_range)
            area = random.uniform(max(1.0, self.area_range[0]), self.area_range[1])
            nw, nh = int(round(w * np.sqrt(area * ratio))), int(round(h * np.sqrt(area / ratio)))
            if nw < w or nh < h:
                continue
            x0, y0 = random.randint(0, nw - w), random.randint(0, nh - h)
            canvas = np.empty((nh, nw, src.shape[2]), dtype=np.float32)
            canvas[:] = np.asarray(self.pad_val, dtype=np.float32)[:src.shape[2]]
            canvas[y0:y0 + h, x0:x0 + w] = src.asnumpy()
            new = label.copy()
            valid = new[:, 0] >= 0
            new[valid, 1] = (new[valid, 1] * w + x0) / nw
            new[valid, 3] = (new[valid, 3] * w + x0) / nw
            new[valid, 2] = (new[valid, 2] * h + y0) / nh
            new[valid, 4] = (new[valid, 4] * h + y0) / nh
            return array(canvas), new
        return src, label


def CreateMultiRandCropAugmenter(min_object_covered=0.1, aspect_ratio_range=(0.75, 1.33), area_range=(0.05, 1.0),
                                 min_eject_coverage=0.3, max_attempts=50, skip_prob=0):
    def align(p):
        return list(p) if isinstance(p, (list, tuple)) and isinstance(p[0], (list, tuple)) or not isinstance(p, (list, tuple)) and False else [p]
    params = [min_object_covered, aspect_ratio_range, area_range, min_eject_coverage, max_attempts]
    n = max(len(p) if isinstance(p, list) else 1 for p in params)
    rows = [p if isinstance(p, list) else [p] * n for p in params]
    augs = [DetRandomCropAug(moc, arr, ar, mec, ma) for moc, arr, ar, mec, ma in zip(*rows)]
    return DetRandomSelectAug(augs, skip_prob=skip_prob)


def CreateDetAugmenter(data_shape, resize=0, rand_crop=0, rand_pad=0, rand_gray=0, rand_mirror=False, mean=None, std=None,
                       brightness=0, contrast=0, saturation=0, pca_noise=0, hue=0, inter_method=2, min_object_covered=0.1,
                       aspect_ratio_range=(0.75, 1.33), area_range=(0.05, 3.0), min_eject_coverage=0.3, max_attempts=50,
                       pad_val=(127, 127, 127)):
    auglist = []
    if resize > 0:
        auglist.append(DetBorrowAug(_img.ResizeAug(resize, inter_method)))
    if rand_crop > 0:
        auglist.append(CreateMultiRandCropAugmenter(min_object_covered, aspect_ratio_range, (area_range[0], min(1.0, area_range[1])),
                                                    min_eject_coverage, max_attempts, skip_prob=1 - rand_crop))
    if rand_mirror > 0:
        auglist.append(DetHorizontalFlipAug(0.5))
    if rand_pad > 0:
        auglist.append(DetRandomSelectAug([DetRandomPadAug(aspect_ratio_range, (1.0, area_range[1]), max_attempts, pad_val)],
                                          1 - rand_pad))
    auglist.append(DetBorrowAug(_img.ForceResizeAug((data_shape[2], data_shape[1]), inter_method)))
    auglist.append(DetBorrowAug(_img.CastAug()))
    if brightness or contrast or saturation:
        auglist.append(DetBorrowAug(_img.ColorJitterAug(brightness, contrast, saturation)))
    if hue:
        auglist.append(DetBorrowAug(_img.HueJitterAug(hue)))
    if pca_noise > 0:
        eigval = np.array([55.46, 4.794, 1.148])
        eigvec = np.array([[-0.5675, 0.7192, 0.4009], [-0.5808, -0.0045, -0.8140], [-0.5836, -0.6948, 0.4203]])
        auglist.append(DetBorrowAug(_img.LightingAug(pca_noise, eigval, eigvec)))
    if rand_gray > 0:
        auglist.append(DetBorrowAug(_img.RandomGrayAug(rand_gray)))
    if mean is True:
        mean = np.array([123.68, 116.28, 103.53])
    if std is True:
        std = np.array([58.395, 57.12, 57.375])
    if mean is not None or std is not None:
        auglist.append(DetBorrowAug(_img.ColorNormalizeAug(mean, std)))
    return auglist


class ImageDetIter(ImageIter):
    """Detection iterator: labels are variable-length object lists packed as
    [header_width, object_width, (extra header...), obj0..., obj1..., ...] and padded with -1 to a common shape
    (detection.py:624-900)."""

    def __init__(self, batch_size, data_shape, path_imgrec=None, path_imglist=None, path_root=None, path_imgidx=None,
                 shuffle=False, part_index=0, num_parts=1, aug_list=None, imglist=None, data_name="data",
                 label_name="label", last_batch_handle="pad", **kwargs):
        super(ImageDetIter, self).__init__(batch_size=batch_size, data_shape=data_shape, path_imgrec=path_imgrec,
                                           path_imglist=path_imglist, path_root=path_root, path_imgidx=path_imgidx,
                                           shuffle=shuffle, part_index=part_index, num_parts=num_parts, aug_list=[],
                                           imglist=imglist, data_name=data_name, label_name=label_name)
        self.auglist = aug_list if aug_list is not None else CreateDetAugmenter(data_shape, **kwargs)
        self.label_shape = self._estimate_label_shape()
        self.provide_label = [_io.DataDesc(label_name, (batch_size,) + self.label_shape, "float32")]

    @staticmethod
    def _parse_label(label):
        raw = np.asarray(label, dtype=np.float32).ravel()
        if raw.size < 7:
            raise RuntimeError("Label shape is invalid: " + str(raw.shape))
        header_width, obj_width = int(raw[0]), int(raw[1])
        if (raw.size - header_width) % obj_width != 0:
            raise RuntimeError("Label shape %s inconsistent with annotation width %d." % (str(raw.shape), obj_width))
        out = raw[header_width:].reshape(-1, obj_width)
        valid = np.where(np.logical_and(out[:, 3] > out[:, 1], out[:, 4] > out[:, 2]))[0]
        if valid.size < 1:
            raise RuntimeError("Encounter sample with no valid label.")
        return out[valid]

    def _estimate_label_shape(self):
        max_count, width = 0, 5
        self.reset()
        try:
            while True:
                label, _ = ImageIter.next_sample(self)
                lab = self._parse_label(label)
                max_count = max(max_count, lab.shape[0])
                width = lab.shape[1]
        except StopIteration:
            pass
        self.reset()
        return (max(max_count, 1), width)

    def reshape(self, data_shape=None, label_shape=None):
        if data_shape is not None:
            self.data_shape = tuple(data_shape)
            self.provide_data = [_io.DataDesc(self.provide_data[0].name, (self.batch_size,) + self.data_shape, "float32")]
        if label_shape is not None:
            self.label_shape = tuple(label_shape)
            self.provide_label = [_io.DataDesc(self.provide_label[0].name, (self.batch_size,) + self.label_shape, "float32")]

    def next(self):
        bs = self.batch_size
        c, h, w = self.data_shape
        data = np.zeros((bs, c, h, w), dtype=np.float32)
        label = np.full((bs,) + self.label_shape, -1.0, dtype=np.float32)
        i = 0
        try:
            while i < bs:
                lab, img = ImageIter.next_sample(self)
                try:
                    lab = self._parse_label(lab)
                except RuntimeError:
                    continue
                for aug in self.auglist:
                    img, lab = aug(img, lab)
                data[i] = img.asnumpy().transpose(2, 0, 1)
                n = min(lab.shape[0], self.label_shape[0])
                label[i, :n] = lab[:n]
                i += 1
        except StopIteration:
            if i == 0:
                raise
        return _io.DataBatch([array(data)], [array(label)], pad=bs - i)

    def sync_label_shape(self, it, verbose=False):
        assert isinstance(it, ImageDetIter), "Synchronize with invalid iterator."
        shape = (max(self.label_shape[0], it.label_shape[0]), max(self.label_shape[1], it.label_shape[1]))
        self.reshape(None, shape)
        it.reshape(None, shape)
        return it
