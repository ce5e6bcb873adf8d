"""contrib.io.DataLoaderIter: wrap a gluon DataLoader as a Module-style DataIter (reference: python/mxnet/contrib/io.py)."""
from ..io import DataBatch, DataDesc, DataIter

__all__ = ["DataLoaderIter"]


class DataLoaderIter(DataIter):
    def __init__(self, loader, data_name="data", label_name="softmax_label", dtype="float32"):
        super(DataLoaderIter, self).__init__()
        self._loader, self._iter = loader, iter(loader)
        data, label = next(self._iter)
        self.batch_size = data.shape[0]
        self.dtype = dtype
        self.provide_data = [DataDesc(data_name, tuple(data.shape), dtype)]
        self.provide_label = [DataDesc(label_name, tuple(label.shape), dtype)]
        self._current_batch = None
        self.reset()

    def reset(self):
        self._iter = iter(self._loader)

    def iter_next(self):
        try:
            self._current_batch = next(self._iter)
        except StopIteration:
            self._current_batch = None
        return self._current_batch is not None

    def getdata(self):
        return [self._pad(self._current_batch[0]).astype(self.dtype)]

    def getlabel(self):
        return [self._pad(self._current_batch[1]).astype(self.dtype)]

    def _pad(self, arr):
        n = arr.shape[0]
        if n == self.batch_size:
            return arr
        from .. import ndarray as nd
        filler = nd.zeros((self.batch_size - n,) + tuple(arr.shape[1:]), dtype=arr.dtype, ctx=arr.context)
        return nd.concat(arr, filler, dim=0)

    def getpad(self):
        return self.batch_size - self._current_batch[0].shape[0]

    def getindex(self):
        return None

    def next(self):
        if self.iter_next():
            return DataBatch(data=self.getdata(), label=self.getlabel(), pad=self.getpad(), index=self.getindex())
        raise StopIteration
