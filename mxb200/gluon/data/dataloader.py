"""DataLoader (reference: python/mxnet/gluon/data/dataloader.py).  Worker processes of the reference
(multiprocessing + POSIX shared-memory NDArrays, CPUSharedStorageManager) become a thread pool feeding pinned
host batches: decoding work releases the GIL in numpy/torch, and pinned staging lets the H2D copy overlap."""
import concurrent.futures as cf

import numpy as np
import torch

from ...ndarray import NDArray, array
from ... import context
from . import sampler as _sampler

__all__ = ["DataLoader", "default_batchify_fn"]


def default_batchify_fn(data):
    if isinstance(data[0], NDArray):
        return NDArray(torch.stack([d._t for d in data]), data[0].context)
    if isinstance(data[0], tuple):
        data = zip(*data)
        return [default_batchify_fn(i) for i in data]
    data = np.asarray(data)
    return array(data, dtype=data.dtype if data.dtype != np.float64 else np.float32)


class DataLoader(object):
    def __init__(self, dataset, batch_size=None, shuffle=False, sampler=None, last_batch=None, batch_sampler=None,
                 batchify_fn=None, num_workers=0, pin_memory=False):
        self._dataset = dataset
        self._pin_memory = pin_memory
        if batch_sampler is None:
            if batch_size is None:
                raise ValueError("batch_size must be specified unless batch_sampler is specified")
            if sampler is None:
                sampler = _sampler.RandomSampler(len(dataset)) if shuffle else _sampler.SequentialSampler(len(dataset))
            elif shuffle:
                raise ValueError("shuffle must not be specified if sampler is specified")
            batch_sampler = _sampler.BatchSampler(sampler, batch_size, last_batch if last_batch else "keep")
        elif batch_size is not None or shuffle or sampler is not None or last_batch is not None:
            raise ValueError("batch_size, shuffle, sampler and last_batch must not be specified if batch_sampler is specified.")
        self._batch_sampler = batch_sampler
        self._num_workers = num_workers if num_workers >= 0 else 0
        self._batchify_fn = batchify_fn or default_batchify_fn

    def _load(self, batch):
        out = self._batchify_fn([self._dataset[idx] for idx in batch])
        if self._pin_memory and torch.cuda.is_available():
            def pin(x):
                if isinstance(x, NDArray):
                    return NDArray(x._t.pin_memory(), context.cpu_pinned())
                if isinstance(x, (list, tuple)):
                    return [pin(i) for i in x]
                return x
            out = pin(out)
        return out

    def __iter__(self):
        if self._num_workers == 0:
            for batch in self._batch_sampler:
                yield self._load(batch)
            return
        with cf.ThreadPoolExecutor(self._num_workers) as ex:
            pending = []
            it = iter(self._batch_sampler)
            for _ in range(2 * self._num_workers):
                b = next(it, None)
                if b is None:
                    break
                pending.append(ex.submit(self._load, b))
            while pending:
                fut = pending.pop(0)
                b = next(it, None)
                if b is not None:
                    pending.append(ex.submit(self._load, b))
                yield fut.result()

    def __len__(self):
        return len(self._batch_sampler)
