"""``mx.sym`` namespace: every op registered in ``mx.nd`` is available symbolically."""
import sys as _sys
import types as _types

from .symbol import Symbol, var, Variable, Group, load, load_json, zeros, ones, _create  # noqa: F401
from ..ndarray import op as _ndop


def _lookup_op(name):
    if _ndop.get(name) is None:
        return None

    def f(*args, **kwargs):
        return _create(name, args, kwargs)
    f.__name__ = name
    return f


def __getattr__(name):
    f = _lookup_op(name)
    if f is None:
        raise AttributeError("module 'mxb200.symbol' has no attribute %r" % name)
    globals()[name] = f
    return f


class _Contrib(_types.ModuleType):
    def __getattr__(self, name):
        f = _lookup_op("_contrib_" + name)
        if f is None:
            raise AttributeError(name)
        return f


contrib = _Contrib(__name__ + ".contrib")
_sys.modules[contrib.__name__] = contrib
