"""BaseDataIterator: user hook that (re)builds the data iterators for the current worker membership
(reference: python/mxnet/module/base_data_iterator.py:1-3; example ETDataIterator / ElasticImageRecordIter,
example/dynamic-training/train_resnet.py:353-390)."""


class BaseDataIterator(object):
    def get_data_iterator(self, kv):
        """Return (train_iter, val_iter) sharded by ``kv.rank`` / ``kv.num_workers``."""
        raise NotImplementedError()


class ElasticDataIterator(BaseDataIterator):
    """Convenience implementation with the reference's invariants: the GLOBAL batch stays constant, the
    per-worker batch is global/num_workers, shards are (num_parts=num_workers, part_index=rank), iterators are
    cached per (num_parts, part_index) and resized to ``epoch_size`` batches."""

    def __init__(self, make_iter, global_batch_size, num_examples=None, make_val_iter=None):
        self._make_iter, self._make_val = make_iter, make_val_iter
        self.global_batch_size = global_batch_size
        self.num_examples = num_examples
        self._cache = {}

    def get_data_iterator(self, kv):
        from ..io import ResizeIter
        nw, rank = max(1, kv.num_workers), kv.rank
        key = (nw, rank)
        if key not in self._cache:
            batch = self.global_batch_size // nw
            train = self._make_iter(batch_size=batch, num_parts=nw, part_index=rank)
            if self.num_examples:
                train = ResizeIter(train, max(1, self.num_examples // self.global_batch_size))
            val = self._make_val(batch_size=batch, num_parts=nw, part_index=rank) if self._make_val else None
            self._cache[key] = (train, val)
        train, val = self._cache[key]
        train.reset()
        return train, val
