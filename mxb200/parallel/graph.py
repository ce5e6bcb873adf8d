"""Whole-step CUDA-graph capture: forward + loss + backward + fused allreduce/update as ONE graph launch.

Why it exists (B200-first): a ResNet-50 step is ~340 kernel launches dispatched from Python.  On one
GPU the launch queue hides the dispatch latency; on 8 GPUs every rank ends its step in the cross-GPU flag barrier of the
fused allreduce kernel (csrc/cuda/allreduce.cu ``peer_barrier``), so the slowest *host* sets the pace of all eight
devices.  Replaying a captured graph removes the host from the steady-state step: one ``cudaGraphLaunch`` per
iteration, kernels back to back on the device, ranks arrive at the barrier within microseconds of each other.

Reference analogue: ``CachedOp`` with ``static_alloc=True, static_shape=True`` (src/imperative/cached_op.cc:560-760,
``HybridBlock.hybridize(static_alloc=..., static_shape=...)``, python/mxnet/gluon/block.py:504-523) -- pre-planned
memory and bulked engine pushes for a fixed-shape forward/backward; the optimizer step stays outside there, here it is
inside because it is a kernel of ours that reads its per-step scalars from device memory.

Contract of :class:`TrainStep`:
  * inputs of one (shape, dtype) signature are copied into static device buffers, then the graph is replayed;
  * the first ``warmup`` calls per signature run eagerly (lazy trainer fusion, allocator / library settling);
  * per-step host scalars (learning-rate schedule, rescale, barrier sequence numbers, Adam/LAMB bias corrections) are
    uploaded into the trainer's hyper buffer before every replay (``DeviceTrainer.step(phase='host')``) -- a captured
    kernel launch would otherwise freeze them;
  * the returned loss NDArray aliases a static buffer: consume (or copy) it before the next call.
"""
import torch

from .. import autograd
from ..base import getenv
from ..ndarray import NDArray

__all__ = ["TrainStep"]


class _Captured(object):
    __slots__ = ("graph", "x", "y", "loss", "calls", "kernels")

    def __init__(self):
        self.graph = None
        self.x = self.y = self.loss = None
        self.calls = 0
        self.kernels = 0


class TrainStep(object):
    """``loss = step(x, y)`` == ``with record(): loss = loss_fn(net(x), y); loss.backward(); trainer.step(batch)``.

    capture=None: capture on CUDA when the trainer runs the fused (arena) path, eager otherwise.
    """

    def __init__(self, net, loss_fn, trainer, batch_size=None, capture=None, warmup=3):
        self.net, self.loss_fn, self.trainer = net, loss_fn, trainer
        self.batch_size = batch_size
        self.capture = capture if capture is not None else bool(getenv("MXB200_CUDA_GRAPH", True))
        self.warmup = max(1, int(warmup))
        self._cache = {}
        self._fused_id = None
        self._stream = None            # side stream: warm-up AND capture run on it (never the legacy default stream)
        self.replays = 0
        self.eager_calls = 0

    # ------------------------------------------------------------------ helpers
    def _eager(self, x, y, bs):
        with autograd.record():
            out = self.net(x)
            loss = self.loss_fn(out, y)
        loss.backward()
        self.trainer.step(bs)
        self.eager_calls += 1
        return loss

    def _can_capture(self, x):
        t = x.as_torch() if isinstance(x, NDArray) else x
        if not (self.capture and t.is_cuda):
            return False
        tr = self.trainer
        if not tr._kv_initialized:
            return False
        return tr._fused is not None

    def _side_stream(self, dev):
        if self._stream is None:
            self._stream = torch.cuda.Stream(dev)
        return self._stream

    def _eager_side(self, x, y, bs):
        """Warm-up step on the capture stream: library handles (cuBLAS / cuDNN workspaces), per-stream scratch buffers and
        the autograd threads' stream bookkeeping are initialised for THAT stream, so nothing has to touch the legacy
        stream while capturing."""
        dev = x.as_torch().device
        s = self._side_stream(dev)
        cur = torch.cuda.current_stream(dev)
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            loss = self._eager(x, y, bs)
        cur.wait_stream(s)
        return loss

    def _capture(self, key, x, y, bs):
        c = self._cache[key]
        tx, ty = x.as_torch(), y.as_torch()
        c.x = torch.empty_like(tx)
        c.y = torch.empty_like(ty)
        c.x.copy_(tx)
        c.y.copy_(ty)
        ctx = x.context
        sx, sy = NDArray(c.x, ctx), NDArray(c.y, ctx)
        tr = self.trainer
        dev = tx.device
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        # the scalars of THIS step go to the hyper buffer first (the captured launches only read it)
        tr._graph_phase = "host"
        tr.step(bs)
        tr._graph_phase = "device"
        from .. import ops as _ops
        k0 = _ops.launch_count()
        try:
            with torch.cuda.graph(g, stream=self._side_stream(dev), capture_error_mode="thread_local"):
                with autograd.record():
                    out = self.net(sx)
                    loss = self.loss_fn(out, sy)
                loss.backward()
                tr.step(bs)
        finally:
            tr._graph_phase = None
        c.graph = g
        c.loss = loss
        c.kernels = _ops.launch_count() - k0       # kernels of ours recorded in the graph (none executed yet)
        _ops.REPLAYED[0] -= c.kernels               # ... so take the capture-time count back out
        # capture only records: run the step whose scalars were just uploaded
        g.replay()
        _ops.REPLAYED[0] += c.kernels
        self.replays += 1
        return c.loss

    # ------------------------------------------------------------------ call
    def reset(self):
        """Drop every captured graph (the parameter / gradient arenas moved: membership change, re-fused trainer)."""
        self._cache = {}
        self._fused_id = None

    def __call__(self, x, y, batch_size=None):
        bs = batch_size or self.batch_size or x.shape[0]
        if not self._can_capture(x):
            return self._eager(x, y, bs)
        if self._fused_id is not None and self._fused_id != id(self.trainer._fused):
            self.reset()                     # graphs captured over the old arenas must never be replayed
        self._fused_id = id(self.trainer._fused)
        tx, ty = x.as_torch(), y.as_torch()
        key = (tuple(tx.shape), tx.dtype, tuple(ty.shape), ty.dtype, bs)
        c = self._cache.get(key)
        if c is None:
            c = self._cache[key] = _Captured()
        if c.graph is None:
            c.calls += 1
            if c.calls <= self.warmup:
                return self._eager_side(x, y, bs)
            return self._capture(key, x, y, bs)
        if c.x.data_ptr() != tx.data_ptr():
            c.x.copy_(tx)
        if c.y.data_ptr() != ty.data_ptr():
            c.y.copy_(ty)
        tr = self.trainer
        tr._graph_phase = "host"
        try:
            tr.step(bs)
        finally:
            tr._graph_phase = None
        c.graph.replay()
        from .. import ops as _ops
        _ops.REPLAYED[0] += c.kernels
        self.replays += 1
        return c.loss

    def static_inputs(self, x, y, batch_size=None):
        """The static (x, y) buffers of the graph captured for this signature, or None.  An input pipeline that writes
        the next batch straight into them (``mx.io.DevicePrefetcher(out=...)``) saves the device-to-device copy."""
        bs = batch_size or self.batch_size or x.shape[0]
        tx, ty = x.as_torch(), y.as_torch()
        c = self._cache.get((tuple(tx.shape), tx.dtype, tuple(ty.shape), ty.dtype, bs))
        if c is None or c.graph is None:
            return None
        return c.x, c.y
