"""Device contexts (reference: python/mxnet/context.py, include/mxnet/base.h:136-191).

``mx.cpu(i)`` contexts with different ids are *distinct logical devices* that
all live in host memory -- the reference's own tests simulate multi-device
kvstores that way (tests/python/unittest/test_kvstore.py) and so do ours.

Provenance: adapted from Apache MXNet (incubating) 1.3.1 ``python/mxnet/context.py`` (Apache License 2.0).  The public API, argument
names and control flow follow the original so that user scripts keep working; the tensor maths runs on torch /
the mxb200 kernels.  This file is API-compatibility glue, not original B200 work ("Provenance").
"""
import threading

import torch

__all__ = ["Context", "cpu", "gpu", "cpu_pinned", "current_context", "num_gpus"]


class _ContextMeta(type):
    """``Context.default_ctx`` (deprecated in the reference but still used): per-thread default context, readable and
    assignable on the class."""

    @property
    def default_ctx(cls):
        return current_context()

    @default_ctx.setter
    def default_ctx(cls, ctx):
        Context._default.ctx = ctx


class Context(object, metaclass=_ContextMeta):
    devtype2str = {1: "cpu", 2: "gpu", 3: "cpu_pinned", 5: "cpu_shared"}
    devstr2type = {"cpu": 1, "gpu": 2, "cpu_pinned": 3, "cpu_shared": 5}
    _default = threading.local()

    def __init__(self, device_type, device_id=0):
        if isinstance(device_type, Context):
            self.device_typeid = device_type.device_typeid
            self.device_id = device_type.device_id
        else:
            self.device_typeid = Context.devstr2type[device_type]
            self.device_id = device_id
        self._old_ctx = None

    @property
    def device_type(self):
        return Context.devtype2str[self.device_typeid]

    @property
    def torch_device(self):
        if self.device_typeid == 2:
            return torch.device("cuda", self.device_id)
        return torch.device("cpu")

    def __hash__(self):
        return hash((self.device_typeid, self.device_id))

    def __eq__(self, other):
        return isinstance(other, Context) and self.device_typeid == other.device_typeid \
            and self.device_id == other.device_id

    def __str__(self):
        return "%s(%d)" % (self.device_type, self.device_id)

    __repr__ = __str__

    def __enter__(self):
        self._old_ctx = getattr(Context._default, "ctx", None)
        Context._default.ctx = self
        return self

    def __exit__(self, *a):
        Context._default.ctx = self._old_ctx


def cpu(device_id=0):
    return Context("cpu", device_id)


def gpu(device_id=0):
    return Context("gpu", device_id)


def cpu_pinned(device_id=0):
    return Context("cpu_pinned", device_id)


def current_context():
    ctx = getattr(Context._default, "ctx", None)
    return ctx if ctx is not None else Context("cpu", 0)


def num_gpus():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def from_torch_device(dev):
    dev = torch.device(dev)
    if dev.type == "cuda":
        return Context("gpu", dev.index or 0)
    return Context("cpu", 0)
