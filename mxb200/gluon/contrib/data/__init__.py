"""gluon.contrib.data (reference: python/mxnet/gluon/contrib/data/{sampler,text}.py)."""
from ...data import sampler as _sampler
from . import text  # noqa: F401

__all__ = ["IntervalSampler", "text"]


class IntervalSampler(_sampler.Sampler):
    """Samples i, i+interval, i+2*interval, ... for i = 0 (and, with ``rollover``, for i = 1 .. interval-1 afterwards)."""

    def __init__(self, length, interval, rollover=True):
        assert interval < length, "Interval {} must be smaller than length {}".format(interval, length)
        self._length, self._interval, self._rollover = length, interval, rollover

    def __iter__(self):
        for i in range(self._interval if self._rollover else 1):
            for j in range(i, self._length, self._interval):
                yield j

    def __len__(self):
        return self._length if self._rollover else len(range(0, self._length, self._interval))
