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
