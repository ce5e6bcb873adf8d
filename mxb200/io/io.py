"""Data iterators (reference: python/mxnet/io/io.py; C++ iterators src/io/*).

``NDArrayIter`` / ``ResizeIter`` / ``PrefetchingIter`` follow the reference semantics (padding, roll-over,
discard; ResizeIter is what the elastic example wraps around each shard, io.py:280).  ``num_parts`` /
``part_index`` sharding is supported by every file-backed iterator, because elastic training re-creates the
iterators with a new (num_parts, part_index) whenever the worker set changes (train_resnet.py:353-377).
"""
import threading
from collections import namedtuple, OrderedDict

import numpy as np
import torch

from ..base import MXNetError
from ..ndarray import NDArray, array
from ..ndarray import sparse as _sp

__all__ = ["DataDesc", "DataBatch", "DataIter", "ResizeIter", "PrefetchingIter", "NDArrayIter", "MXDataIter",
           "CSVIter", "MNISTIter", "LibSVMIter", "ImageRecordIter", "SyntheticDataIter"]


class DataDesc(namedtuple("DataDesc", ["name", "shape"])):
    def __new__(cls, name, shape, dtype=np.float32, layout="NCHW"):
        ret = super(cls, DataDesc).__new__(cls, name, tuple(shape))
        ret.dtype = dtype
        ret.layout = layout
        return ret

    def __repr__(self):
        return "DataDesc[%s,%s,%s,%s]" % (self.name, self.shape, self.dtype, self.layout)

    @staticmethod
    def get_batch_axis(layout):
        if layout is None:
            return 0
        return layout.find("N")

    @staticmethod
    def get_list(shapes, types):
        if types is not None:
            type_dict = dict(types)
            return [DataDesc(x[0], x[1], type_dict[x[0]]) for x in shapes]
        return [DataDesc(x[0], x[1]) for x in shapes]


class DataBatch(object):
    def __init__(self, data, label=None, pad=None, index=None, bucket_key=None, provide_data=None, provide_label=None):
        if data is not None:
            assert isinstance(data, (list, tuple)), "Data must be list of NDArrays"
        if label is not None:
            assert isinstance(label, (list, tuple)), "Label must be list of NDArrays"
        self.data = data
        self.label = label
        self.pad = pad
        self.index = index
        self.bucket_key = bucket_key
        self.provide_data = provide_data
        self.provide_label = provide_label

    def __str__(self):
        data_shapes = [d.shape for d in self.data]
        label_shapes = [l.shape for l in self.label] if self.label else None
        return "{}: data shapes: {} label shapes: {}".format(self.__class__.__name__, data_shapes, label_shapes)


class DataIter(object):
    def __init__(self, batch_size=0):
        self.batch_size = batch_size

    def __iter__(self):
        return self

    def reset(self):
        pass

    def next(self):
        if self.iter_next():
            return DataBatch(data=self.getdata(), label=self.getlabel(), pad=self.getpad(), index=self.getindex())
        raise StopIteration

    def __next__(self):
        return self.next()

    def iter_next(self):
        pass

    def getdata(self):
        pass

    def getlabel(self):
        pass

    def getindex(self):
        return None

    def getpad(self):
        pass


class ResizeIter(DataIter):
    """Resize an iterator to ``size`` batches per epoch (wraps around / truncates)."""

    def __init__(self, data_iter, size, reset_internal=True):
        super(ResizeIter, self).__init__()
        self.data_iter = data_iter
        self.size = size
        self.reset_internal = reset_internal
        self.cur = 0
        self.current_batch = None
        self.provide_data = data_iter.provide_data
        self.provide_label = data_iter.provide_label
        self.batch_size = data_iter.batch_size
        if hasattr(data_iter, "default_bucket_key"):
            self.default_bucket_key = data_iter.default_bucket_key

    def reset(self):
        self.cur = 0
        if self.reset_internal:
            self.data_iter.reset()

    def iter_next(self):
        if self.cur == self.size:
            return False
        try:
            self.current_batch = self.data_iter.next()
        except StopIteration:
            self.data_iter.reset()
            self.current_batch = self.data_iter.next()
        self.cur += 1
        return True

    def getdata(self):
        return self.current_batch.data

    def getlabel(self):
        return self.current_batch.label

    def getindex(self):
        return self.current_batch.index

    def getpad(self):
        return self.current_batch.pad


class PrefetchingIter(DataIter):
    """Background-thread prefetch of one or more iterators (io.py:345-470)."""

    def __init__(self, iters, rename_data=None, rename_label=None):
        super(PrefetchingIter, self).__init__()
        if not isinstance(iters, list):
            iters = [iters]
        self.n_iter = len(iters)
        assert self.n_iter > 0
        self.iters = iters
        self.rename_data = rename_data
        self.rename_label = rename_label
        self.batch_size = self.provide_data[0][1][0]
        self.data_ready = [threading.Event() for _ in range(self.n_iter)]
        self.data_taken = [threading.Event() for _ in range(self.n_iter)]
        for e in self.data_taken:
            e.set()
        self.started = True
        self.current_batch = [None for _ in range(self.n_iter)]
        self.next_batch = [None for _ in range(self.n_iter)]

        def prefetch_func(self, i):
            while True:
                self.data_taken[i].wait()
                if not self.started:
                    break
                try:
                    self.next_batch[i] = self.iters[i].next()
                except StopIteration:
                    self.next_batch[i] = None
                self.data_taken[i].clear()
                self.data_ready[i].set()
        self.prefetch_threads = [threading.Thread(target=prefetch_func, args=[self, i]) for i in range(self.n_iter)]
        for thread in self.prefetch_threads:
            thread.daemon = True
            thread.start()

    def __del__(self):
        self.started = False
        for e in self.data_taken:
            e.set()
        for thread in self.prefetch_threads:
            thread.join(timeout=1.0)

    @property
    def provide_data(self):
        if self.rename_data is None:
            return sum([i.provide_data for i in self.iters], [])
        return sum([[DataDesc(r[x.name], x.shape, x.dtype) if isinstance(x, DataDesc) else DataDesc(*x) for x in i.provide_data]
                    for r, i in zip(self.rename_data, self.iters)], [])

    @property
    def provide_label(self):
        if self.rename_label is None:
            return sum([i.provide_label for i in self.iters], [])
        return sum([[DataDesc(r[x.name], x.shape, x.dtype) if isinstance(x, DataDesc) else DataDesc(*x) for x in i.provide_label]
                    for r, i in zip(self.rename_label, self.iters)], [])

    def reset(self):
        for e in self.data_ready:
            e.wait()
        for i in self.iters:
            i.reset()
        for e in self.data_ready:
            e.clear()
        for e in self.data_taken:
            e.set()

    def iter_next(self):
        for e in self.data_ready:
            e.wait()
        if self.next_batch[0] is None:
            for i in self.next_batch:
                assert i is None, "Number of entry mismatches between iterators"
            return False
        for batch in self.next_batch:
            assert batch.pad == self.next_batch[0].pad, "Number of entry mismatches between iterators"
        self.current_batch = DataBatch(sum([batch.data for batch in self.next_batch], []),
                                       sum([batch.label for batch in self.next_batch], []),
                                       self.next_batch[0].pad, self.next_batch[0].index,
                                       provide_data=self.provide_data, provide_label=self.provide_label)
        for e in self.data_ready:
            e.clear()
        for e in self.data_taken:
            e.set()
        return True

    def next(self):
        if self.iter_next():
            return self.current_batch
        raise StopIteration

    def getdata(self):
        return self.current_batch.data

    def getlabel(self):
        return self.current_batch.label

    def getindex(self):
        return self.current_batch.index

    def getpad(self):
        return self.current_batch.pad


def _init_data(data, allow_empty, default_name):
    assert (data is not None) or allow_empty
    if data is None:
        data = []
    if isinstance(data, (np.ndarray, NDArray, torch.Tensor)) or hasattr(data, "stype") \
            or type(data).__module__.startswith("scipy.sparse"):
        data = [data]
    if isinstance(data, list):
        if not allow_empty:
            assert len(data) > 0
        if len(data) == 1:
            data = OrderedDict([(default_name, data[0])])
        else:
            data = OrderedDict([("_%d_%s" % (i, default_name), d) for i, d in enumerate(data)])
    if not isinstance(data, dict):
        raise TypeError("Input must be NDArray, numpy.ndarray, a list of them or dict with them as values")
    out = []
    for k, v in data.items():
        if type(v).__module__.startswith("scipy.sparse"):
            v = _sp.csr_matrix(v.tocsr(), dtype=v.dtype if v.dtype != np.float64 else np.float32)    # scipy -> CSRNDArray
        if not isinstance(v, (NDArray, _sp.BaseSparseNDArray)):
            try:
                v = array(v, dtype=v.dtype if hasattr(v, "dtype") and not isinstance(v, torch.Tensor) and v.dtype != np.float64 else None)
            except Exception:
                raise TypeError(("Invalid type '%s' for %s, " % (type(v), k)) + "should be NDArray, numpy.ndarray")
        out.append((k, v))
    return out


class NDArrayIter(DataIter):
    """Iterate over in-memory arrays with shuffle and last_batch_handle in {'pad','discard','roll_over'}."""

    def __init__(self, data, label=None, batch_size=1, shuffle=False, last_batch_handle="pad", data_name="data",
                 label_name="softmax_label"):
        super(NDArrayIter, self).__init__(batch_size)
        self.data = _init_data(data, allow_empty=False, default_name=data_name)
        self.label = _init_data(label, allow_empty=True, default_name=label_name)
        if any(isinstance(v, _sp.CSRNDArray) for _, v in self.data + self.label) and last_batch_handle != "discard":
            raise NotImplementedError("`NDArrayIter` only supports ``CSRNDArray`` with `last_batch_handle` set to `discard`.")
        self.idx = np.arange(self.data[0][1].shape[0])
        if shuffle:
            np.random.shuffle(self.idx)
            self.data = [(k, self._take(v, self.idx)) for k, v in self.data]
            self.label = [(k, self._take(v, self.idx)) for k, v in self.label]
        if last_batch_handle == "discard":
            new_n = self.data[0][1].shape[0] - self.data[0][1].shape[0] % batch_size
            self.idx = self.idx[:new_n]
        self.data_list = [x[1] for x in self.data] + [x[1] for x in self.label]
        self.num_source = len(self.data_list)
        self.num_data = self.idx.shape[0]
        assert self.num_data >= batch_size, "batch_size needs to be smaller than data size."
        self.cursor = -batch_size
        self.batch_size = batch_size
        self.last_batch_handle = last_batch_handle

    @staticmethod
    def _take(v, idx):
        if isinstance(v, _sp.CSRNDArray):
            return _sp.cast_storage(NDArray(v.todense()._t[torch.as_tensor(idx)]), "csr")
        return NDArray(v._t[torch.as_tensor(idx)], v.context)

    @property
    def provide_data(self):
        return [DataDesc(k, tuple([self.batch_size] + list(v.shape[1:])), v.dtype) for k, v in self.data]

    @property
    def provide_label(self):
        return [DataDesc(k, tuple([self.batch_size] + list(v.shape[1:])), v.dtype) for k, v in self.label]

    def hard_reset(self):
        self.cursor = -self.batch_size
        self._carry = 0

    def reset(self):
        """'roll_over': the incomplete tail of the finished epoch is not emitted; it becomes the head of the first batch of
        the next epoch (io.py NDArrayIter of the reference)."""
        carry = getattr(self, "_carry", 0)
        self._carry = 0
        self.cursor = -self.batch_size - carry if self.last_batch_handle == "roll_over" else -self.batch_size

    def iter_next(self):
        self.cursor += self.batch_size
        return self.cursor < self.num_data

    def next(self):
        if self.iter_next():
            if self.last_batch_handle == "roll_over" and self.cursor + self.batch_size > self.num_data:
                self._carry = self.num_data - self.cursor       # keep the tail for the next epoch
                raise StopIteration
            return DataBatch(data=self.getdata(), label=self.getlabel(), pad=self.getpad(), index=None)
        raise StopIteration

    def _getdata(self, data_source):
        assert self.cursor < self.num_data, "DataIter needs reset."
        if self.cursor < 0:          # first batch after a roll-over: tail of the previous epoch + head of this one
            return [NDArray(torch.cat([x._t[self.cursor:], x._t[:self.cursor + self.batch_size]]), x.context)
                    for _, x in data_source]
        if self.cursor + self.batch_size <= self.num_data:
            out = []
            for _, x in data_source:
                if isinstance(x, _sp.CSRNDArray):
                    out.append(x[self.cursor:self.cursor + self.batch_size])
                else:
                    out.append(NDArray(x._t[self.cursor:self.cursor + self.batch_size], x.context))
            return out
        pad = self.batch_size - self.num_data + self.cursor
        return [NDArray(torch.cat([x._t[self.cursor:], x._t[:pad]]), x.context) for _, x in data_source]

    def getdata(self):
        return self._getdata(self.data)

    def getlabel(self):
        return self._getdata(self.label)

    def getpad(self):
        if self.last_batch_handle == "pad" and self.cursor + self.batch_size > self.num_data:
            return self.cursor + self.batch_size - self.num_data
        return 0


class _ArrayBackedCIter(NDArrayIter):
    """NDArrayIter behind the iterators that are C++ classes in the reference (CSVIter, LibSVMIter): with those,
    ``getdata()`` / ``getlabel()`` return THE array of the current batch instead of a list."""

    def next(self):
        if self.iter_next():
            return DataBatch(data=self._getdata(self.data), label=self._getdata(self.label), pad=self.getpad(), index=None)
        raise StopIteration

    def getdata(self):
        return self._getdata(self.data)[0]

    def getlabel(self):
        lab = self._getdata(self.label)
        return lab[0] if lab else None


class SyntheticDataIter(DataIter):
    """One fixed random batch served ``epoch_size`` times per epoch (the reference's --benchmark 1 iterator,
    example/image-classification/common/data.py:98-130)."""

    def __init__(self, num_classes, data_shape, epoch_size=100, dtype="float32", ctx=None, label_dtype="float32",
                 data_name="data", label_name="softmax_label"):
        super(SyntheticDataIter, self).__init__(data_shape[0])
        self.cur_iter = 0
        self.max_iter = epoch_size
        self.dtype = dtype
        label = np.random.randint(0, num_classes, [self.batch_size, ])
        data = np.random.uniform(-1, 1, data_shape)
        self.data = array(data.astype("float32"), ctx=ctx, dtype=dtype)
        self.label = array(label.astype("float32"), ctx=ctx, dtype=label_dtype)
        try:
            import torch
            if torch.cuda.is_available():   # the reference allocates the batch in cpu_pinned memory (common/data.py:113-118)
                for a in (self.data, self.label):
                    if not a._t.is_cuda:
                        a._t = a._t.pin_memory()
        except Exception:  # noqa: BLE001
            pass
        self.provide_data = [DataDesc(data_name, tuple(data_shape), dtype)]
        self.provide_label = [DataDesc(label_name, (self.batch_size,), label_dtype)]

    def next(self):
        self.cur_iter += 1
        if self.cur_iter <= self.max_iter:
            return DataBatch(data=[self.data], label=[self.label], pad=0, index=None, provide_data=self.provide_data,
                             provide_label=self.provide_label)
        raise StopIteration

    def reset(self):
        self.cur_iter = 0


class MXDataIter(DataIter):
    """Adapter base for the native (C++) iterators: a python generator of DataBatch with provide_* metadata."""

    def __init__(self, gen_factory, provide_data, provide_label, batch_size):
        super(MXDataIter, self).__init__(batch_size)
        self._factory = gen_factory
        self._gen = gen_factory()
        self.provide_data, self.provide_label = provide_data, provide_label

    def reset(self):
        self._gen = self._factory()
        self._cur = None

    def next(self):
        self._cur = next(self._gen)
        return self._cur

    # iter_next / getdata / getlabel / getpad protocol of the C++ iterators (io.py MXDataIter of the reference: getdata()
    # is THE data array of the current batch, not a list)
    def iter_next(self):
        try:
            self._cur = next(self._gen)
            return True
        except StopIteration:
            return False

    def getdata(self):
        return self._cur.data[0]

    def getlabel(self):
        return self._cur.label[0] if self._cur.label else None

    def getindex(self):
        return self._cur.index

    def getpad(self):
        return self._cur.pad or 0


def _shard(n, num_parts, part_index):
    per = n // num_parts
    return part_index * per, (part_index + 1) * per if part_index < num_parts - 1 else n if num_parts == 1 else (part_index + 1) * per


def CSVIter(data_csv, data_shape, label_csv=None, label_shape=(1,), batch_size=1, round_batch=True, dtype="float32",
            num_parts=1, part_index=0, **kwargs):
    """CSV iterator (src/io/iter_csv.cc) with num_parts/part_index sharding."""
    data = np.loadtxt(data_csv, delimiter=",", dtype=np.dtype(dtype), ndmin=2).reshape((-1,) + tuple(data_shape))
    label = np.loadtxt(label_csv, delimiter=",", dtype=np.float32, ndmin=2).reshape((-1,) + tuple(label_shape)) \
        if label_csv else np.zeros((data.shape[0],) + tuple(label_shape), np.float32)
    b, e = _shard(len(data), num_parts, part_index)
    if tuple(label_shape) == (1,):
        label = label.reshape(-1)
    return _ArrayBackedCIter(data[b:e], label[b:e], batch_size, last_batch_handle="pad" if round_batch else "discard")


def MNISTIter(image, label, batch_size=128, shuffle=True, flat=False, seed=0, silent=False, num_parts=1, part_index=0,
              **kwargs):
    """MNIST idx-file iterator (src/io/iter_mnist.cc:260)."""
    import gzip
    import struct
    op = gzip.open if image.endswith(".gz") else open
    with op(label, "rb") as f:
        struct.unpack(">II", f.read(8))
        lab = np.frombuffer(f.read(), dtype=np.uint8).astype(np.float32)
    with op(image, "rb") as f:
        _, n, r, c = struct.unpack(">IIII", f.read(16))
        img = np.frombuffer(f.read(), dtype=np.uint8).reshape(n, 1, r, c).astype(np.float32) / 255.0
    if flat:
        img = img.reshape(n, -1)
    b, e = _shard(n, num_parts, part_index)
    img, lab = img[b:e], lab[b:e]
    if shuffle:
        rng = np.random.RandomState(seed)
        p = rng.permutation(len(lab))
        img, lab = img[p], lab[p]
    return NDArrayIter(img, lab, batch_size, shuffle=False, last_batch_handle="discard")


def LibSVMIter(data_libsvm, data_shape, label_libsvm=None, label_shape=(1,), batch_size=1, num_parts=1, part_index=0,
               **kwargs):
    """LibSVM text iterator producing CSR batches (src/io/iter_libsvm.cc)."""
    rows, cols, vals, labels, indptr = [], [], [], [], [0]
    with open(data_libsvm) as f:
        for line in f:
            parts = line.strip().split()
            if not parts:
                continue
            labels.append(float(parts[0]))
            for kv in parts[1:]:
                k, v = kv.split(":")
                cols.append(int(k))
                vals.append(float(v))
            indptr.append(len(cols))
    n = len(labels)
    b, e = _shard(n, num_parts, part_index)
    csr = _sp.csr_matrix((np.array(vals[indptr[b]:indptr[e]], np.float32), np.array(cols[indptr[b]:indptr[e]], np.int64),
                          np.array(indptr[b:e + 1], np.int64) - indptr[b]), shape=(e - b, data_shape[0]))
    return _ArrayBackedCIter(csr, np.array(labels[b:e], np.float32), batch_size, last_batch_handle="discard")


def ImageRecordIter(path_imgrec, data_shape, batch_size, path_imgidx=None, label_width=1, shuffle=False, num_parts=1,
                    part_index=0, rand_crop=False, rand_mirror=False, mean_r=0.0, mean_g=0.0, mean_b=0.0, std_r=1.0,
                    std_g=1.0, std_b=1.0, resize=0, preprocess_threads=4, prefetch_buffer=4, dtype="float32",
                    data_name="data", label_name="softmax_label", ctx=None, layout="NCHW", random_resized_crop=False,
                    min_random_area=1.0, max_random_area=1.0, min_aspect_ratio=None, max_aspect_ratio=0.0, seed=0,
                    round_batch=True, **kwargs):
    """RecordIO image pipeline (src/io/iter_image_recordio_2.cc).

    ``ctx=mx.gpu(i)`` (or ``MXB200_GPU_DECODE=1`` with a visible GPU) selects the device pipeline
    (``io/gpu_pipeline.py``: C++ loader threads, nvJPEG batched decode, fused augmentation kernel; batches are born on
    the GPU in ``layout`` / ``dtype``).  Otherwise batches are produced in host memory by the threaded OpenCV pipeline of
    ``io/native.py`` exactly like the reference's iterator."""
    from ..base import getenv
    use_gpu = ctx is not None and getattr(ctx, "device_type", "cpu") == "gpu"
    if ctx is None and getenv("MXB200_GPU_DECODE", False):
        import torch
        if torch.cuda.is_available():
            from ..context import gpu
            ctx, use_gpu = gpu(0), True
    if use_gpu:
        from .gpu_pipeline import GPUImageRecordIter
        return GPUImageRecordIter(path_imgrec, data_shape, batch_size, path_imgidx=path_imgidx, label_width=label_width,
                                  shuffle=shuffle, num_parts=num_parts, part_index=part_index, rand_crop=rand_crop,
                                  rand_mirror=rand_mirror, random_resized_crop=random_resized_crop,
                                  min_random_area=min_random_area, max_random_area=max_random_area,
                                  min_aspect_ratio=min_aspect_ratio, max_aspect_ratio=max_aspect_ratio,
                                  mean=(mean_r, mean_g, mean_b), std=(std_r, std_g, std_b), resize=resize,
                                  threads=preprocess_threads, prefetch=prefetch_buffer, dtype=dtype, layout=layout, ctx=ctx,
                                  seed=seed, round_batch=round_batch, data_name=data_name, label_name=label_name)
    from .native import ImageRecordIterImpl
    return ImageRecordIterImpl(path_imgrec, data_shape, batch_size, path_imgidx=path_imgidx, label_width=label_width,
                               shuffle=shuffle, num_parts=num_parts, part_index=part_index, rand_crop=rand_crop,
                               rand_mirror=rand_mirror, mean=(mean_r, mean_g, mean_b), std=(std_r, std_g, std_b),
                               resize=resize, threads=preprocess_threads, prefetch=prefetch_buffer, dtype=dtype,
                               data_name=data_name, label_name=label_name)
