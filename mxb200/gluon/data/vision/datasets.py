"""MNIST / FashionMNIST / CIFAR readers (local files only: no network) + synthetic stand-ins of the same
shape for tests and benchmarks (reference: python/mxnet/gluon/data/vision/datasets.py)."""
import gzip
import os
import struct

import numpy as np

from ..dataset import Dataset, ArrayDataset
from ....ndarray import array

__all__ = ["MNIST", "FashionMNIST", "CIFAR10", "CIFAR100", "SyntheticImageDataset", "ImageRecordDataset",
           "ImageFolderDataset"]


class SyntheticImageDataset(Dataset):
    """Deterministic random images/labels of a given shape (the reference's --benchmark synthetic mode)."""

    def __init__(self, length, shape=(28, 28, 1), classes=10, dtype="uint8", seed=0):
        rng = np.random.RandomState(seed)
        self._x = rng.randint(0, 256, size=(length,) + tuple(shape)).astype(dtype)
        self._y = rng.randint(0, classes, size=(length,)).astype("int32")

    def __getitem__(self, idx):
        return array(self._x[idx], dtype=self._x.dtype), self._y[idx]

    def __len__(self):
        return len(self._y)


class _Downloaded(Dataset):
    def __init__(self, root, transform):
        self._transform = transform
        self._root = os.path.expanduser(root)
        self._data = self._label = None
        self._get_data()

    def __getitem__(self, idx):
        if self._transform is not None:
            return self._transform(self._data[idx], self._label[idx])
        return self._data[idx], self._label[idx]

    def __len__(self):
        return len(self._label)


class MNIST(_Downloaded):
    _files = {True: ("train-images-idx3-ubyte.gz", "train-labels-idx1-ubyte.gz"),
              False: ("t10k-images-idx3-ubyte.gz", "t10k-labels-idx1-ubyte.gz")}

    def __init__(self, root=os.path.join("~", ".mxnet", "datasets", "mnist"), train=True, transform=None):
        self._train = train
        super(MNIST, self).__init__(root, transform)

    def _get_data(self):
        img, lab = [os.path.join(self._root, f) for f in self._files[self._train]]
        if not (os.path.exists(img) and os.path.exists(lab)):
            raise RuntimeError("MNIST files not found under %s and there is no network to download them; use "
                               "SyntheticImageDataset for shape-compatible data" % self._root)
        with gzip.open(lab, "rb") as fin:
            struct.unpack(">II", fin.read(8))
            label = np.frombuffer(fin.read(), dtype=np.uint8).astype(np.int32)
        with gzip.open(img, "rb") as fin:
            struct.unpack(">IIII", fin.read(16))
            data = np.frombuffer(fin.read(), dtype=np.uint8).reshape(len(label), 28, 28, 1)
        self._data = array(data, dtype=data.dtype)
        self._label = label


class FashionMNIST(MNIST):
    def __init__(self, root=os.path.join("~", ".mxnet", "datasets", "fashion-mnist"), train=True, transform=None):
        super(FashionMNIST, self).__init__(root, train, transform)


class CIFAR10(_Downloaded):
    def __init__(self, root=os.path.join("~", ".mxnet", "datasets", "cifar10"), train=True, transform=None):
        self._train = train
        super(CIFAR10, self).__init__(root, transform)

    def _read_batch(self, filename):
        with open(filename, "rb") as fin:
            data = np.frombuffer(fin.read(), dtype=np.uint8).reshape(-1, 3072 + 1)
        return data[:, 1:].reshape(-1, 3, 32, 32).transpose(0, 2, 3, 1), data[:, 0].astype(np.int32)

    def _get_data(self):
        names = ["data_batch_%d.bin" % i for i in range(1, 6)] if self._train else ["test_batch.bin"]
        files = [os.path.join(self._root, n) for n in names]
        if not all(os.path.exists(f) for f in files):
            raise RuntimeError("CIFAR files not found under %s and there is no network to download them" % self._root)
        data, label = zip(*(self._read_batch(f) for f in files))
        self._data = array(np.concatenate(data), dtype=np.uint8)
        self._label = np.concatenate(label)


class CIFAR100(CIFAR10):
    def __init__(self, root=os.path.join("~", ".mxnet", "datasets", "cifar100"), fine_label=False, train=True, transform=None):
        self._fine_label = fine_label
        super(CIFAR100, self).__init__(root, train, transform)

    def _read_batch(self, filename):
        with open(filename, "rb") as fin:
            data = np.frombuffer(fin.read(), dtype=np.uint8).reshape(-1, 3072 + 2)
        return data[:, 2:].reshape(-1, 3, 32, 32).transpose(0, 2, 3, 1), data[:, 0 + self._fine_label].astype(np.int32)

    def _get_data(self):
        names = ["train.bin"] if self._train else ["test.bin"]
        files = [os.path.join(self._root, n) for n in names]
        if not all(os.path.exists(f) for f in files):
            raise RuntimeError("CIFAR100 files not found under %s and there is no network" % self._root)
        data, label = zip(*(self._read_batch(f) for f in files))
        self._data = array(np.concatenate(data), dtype=np.uint8)
        self._label = np.concatenate(label)


class ImageRecordDataset(Dataset):
    """Images + labels from a RecordIO pack (im2rec format)."""

    def __init__(self, filename, flag=1, transform=None):
        from ..dataset import RecordFileDataset
        self._rec = RecordFileDataset(filename)
        self._flag, self._transform = flag, transform

    def __getitem__(self, idx):
        from .... import recordio, image
        header, img = recordio.unpack(self._rec[idx])
        data = image.imdecode(img, self._flag)
        if self._transform is not None:
            return self._transform(data, header.label)
        return data, header.label

    def __len__(self):
        return len(self._rec)


class ImageFolderDataset(Dataset):
    def __init__(self, root, flag=1, transform=None):
        self._root = os.path.expanduser(root)
        self._flag, self._transform = flag, transform
        self._exts = [".jpg", ".jpeg", ".png", ".npy"]
        self.synsets, self.items = [], []
        for folder in sorted(os.listdir(self._root)):
            path = os.path.join(self._root, folder)
            if not os.path.isdir(path):
                continue
            label = len(self.synsets)
            self.synsets.append(folder)
            for filename in sorted(os.listdir(path)):
                if os.path.splitext(filename)[1].lower() in self._exts:
                    self.items.append((os.path.join(path, filename), label))

    def __getitem__(self, idx):
        from .... import image
        img = image.imread(self.items[idx][0], self._flag)
        label = self.items[idx][1]
        if self._transform is not None:
            return self._transform(img, label)
        return img, label

    def __len__(self):
        return len(self.items)
