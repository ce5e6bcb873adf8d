"""pytest plugin: ``import mxnet`` (and every ``mxnet.<sub>``) resolves to the very same module objects as ``mxb200``.

Used by tools/run_reference_tests.py to run the reference's own python unit tests (read-only under /root/reference, never
copied) against this framework as an API-parity probe.  A meta-path finder is needed instead of a plain
``sys.modules['mxnet'] = mxb200``: otherwise ``import mxnet.gluon`` would execute mxb200/gluon a second time under the
other name and isinstance checks between the two copies fail."""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _REPO)
import mxb200  # noqa: E402


class _Alias(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname == "mxnet" or fullname.startswith("mxnet."):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        real = "mxb200" + spec.name[len("mxnet"):]
        parent, _, child = real.rpartition(".")
        try:
            return importlib.import_module(real)
        except ModuleNotFoundError:
            mod = getattr(importlib.import_module(parent), child, None)   # dynamically created namespaces (nd.contrib ...)
            if mod is None:
                raise
            return mod

    def exec_module(self, module):
        pass


sys.meta_path.insert(0, _Alias())
sys.modules["mxnet"] = mxb200
_ref = os.environ.get("MXB200_REFERENCE_TESTS", "/root/reference/tests/python/unittest")
sys.path.insert(0, _ref)

# ---- numpy 1.x / python 2-3.7 idioms of the 2018 test-suite --------------------------------------------------------
# The reference's tests were written against numpy 1.14 and python <= 3.7.  Three of their idioms are gone from the
# interpreter / numpy of this image; they are restored HERE (test harness only, nothing in mxb200 depends on them) so the
# probe measures this framework's API, not numpy's deprecation schedule:
#   * ``np.int`` / ``np.float`` / ``np.bool`` aliases,
#   * ``time.clock``,
#   * indexing an ndarray with a *list* of slices (``a[[slice(..), slice(..)]]``): NDArray.asnumpy() hands out an ndarray
#     subclass that treats such a list as the tuple numpy 1.x treated it as.
import time  # noqa: E402

import numpy as _np  # noqa: E402

for _alias, _target in (("int", int), ("float", float), ("bool", bool), ("object", object), ("str", str)):
    if _alias not in _np.__dict__:
        setattr(_np, _alias, _target)
if not hasattr(time, "clock"):
    time.clock = time.perf_counter


class _LegacyIndexArray(_np.ndarray):
    def __getitem__(self, key):
        if isinstance(key, list) and key and any(isinstance(k, slice) for k in key):
            key = tuple(key)
        r = _np.ndarray.__getitem__(self, key)
        return r.view(_np.ndarray) if isinstance(r, _np.ndarray) else r


_plain_asnumpy = mxb200.nd.NDArray.asnumpy


def _asnumpy(self):
    a = _plain_asnumpy(self)
    return a.view(_LegacyIndexArray) if isinstance(a, _np.ndarray) and a.ndim > 0 else a


mxb200.nd.NDArray.asnumpy = _asnumpy

_plain_printoptions = _np.set_printoptions


def _set_printoptions(*a, **k):          # numpy 1.x accepted threshold=np.nan for "print everything"
    if isinstance(k.get("threshold"), float) and k["threshold"] != k["threshold"]:
        k["threshold"] = sys.maxsize
    return _plain_printoptions(*a, **k)


_np.set_printoptions = _set_printoptions
