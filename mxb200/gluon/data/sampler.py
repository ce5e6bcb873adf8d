"""Samplers (reference: python/mxnet/gluon/data/sampler.py; SplitSampler from
example/distributed_training/cifar10_dist.py:54-78).

Provenance: adapted from Apache MXNet (incubating) 1.3.1 ``python/mxnet/gluon/data/sampler.py`` (Apache License 2.0).  The public API, argument
names and control flow follow the original so that user scripts keep working; the tensor maths runs on torch /
the mxb200 kernels.  This file is API-compatibility glue, not original B200 work ("Provenance").
"""
import random

__all__ = ["Sampler", "SequentialSampler", "RandomSampler", "BatchSampler", "SplitSampler"]


class Sampler(object):
    def __iter__(self):
        raise NotImplementedError

    def __len__(self):
        raise NotImplementedError


class SequentialSampler(Sampler):
    def __init__(self, length):
        self._length = length

    def __iter__(self):
        return iter(range(self._length))

    def __len__(self):
        return self._length


class RandomSampler(Sampler):
    def __init__(self, length):
        self._length = length

    def __iter__(self):
        indices = list(range(self._length))
        random.shuffle(indices)
        return iter(indices)

    def __len__(self):
        return self._length


class SplitSampler(Sampler):
    """Shard ``[0, length)`` into num_parts contiguous parts and sample part_index (elastic data sharding)."""

    def __init__(self, length, num_parts=1, part_index=0, shuffle=True):
        self.part_len = length // num_parts
        self.start = self.part_len * part_index
        self.end = self.start + self.part_len
        self.shuffle = shuffle

    def __iter__(self):
        indices = list(range(self.start, self.end))
        if self.shuffle:
            random.shuffle(indices)
        return iter(indices)

    def __len__(self):
        return self.part_len


class BatchSampler(Sampler):
    def __init__(self, sampler, batch_size, last_batch="keep"):
        self._sampler, self._batch_size, self._last_batch = sampler, batch_size, last_batch
        self._prev = []

    def __iter__(self):
        batch, self._prev = self._prev, []
        for i in self._sampler:
            batch.append(i)
            if len(batch) == self._batch_size:
                yield batch
                batch = []
        if batch:
            if self._last_batch == "keep":
                yield batch
            elif self._last_batch == "discard":
                return
            elif self._last_batch == "rollover":
                self._prev = batch
            else:
                raise ValueError("last_batch must be one of 'keep', 'discard', or 'rollover', but got %s" % self._last_batch)

    def __len__(self):
        if self._last_batch == "keep":
            return (len(self._sampler) + self._batch_size - 1) // self._batch_size
        if self._last_batch == "discard":
            return len(self._sampler) // self._batch_size
        if self._last_batch == "rollover":
            return (len(self._prev) + len(self._sampler)) // self._batch_size
        raise ValueError("last_batch must be one of 'keep', 'discard', or 'rollover', but got %s" % self._last_batch)
