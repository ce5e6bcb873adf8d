"""Automatic naming of symbols (reference: python/mxnet/name.py).

``NameManager`` hands out ``<hint><counter>`` names for operators created without a name; used as a context manager it
scopes the counters, and ``Prefix`` additionally prepends a string to every name created inside the ``with`` block."""
import threading
import warnings

__all__ = ["NameManager", "Prefix"]


class _ManagerMeta(type):
    """``NameManager.current``: the calling thread's manager, readable (deprecated in favour of ``with``) and assignable on
    the class (python/mxnet/name.py of the reference: classproperty over a thread-local)."""

    @property
    def current(cls):
        warnings.warn("NameManager.current has been deprecated. It is advised to use the `with` statement with NameManager.",
                      DeprecationWarning)
        return cls._get_current()

    @current.setter
    def current(cls, value):
        NameManager._current.value = value


class NameManager(object, metaclass=_ManagerMeta):
    _current = threading.local()

    def __init__(self):
        self._counter = {}
        self._old_manager = None

    def get(self, name, hint):
        """The user's name if given, else ``hint`` + a per-hint running number."""
        if name:
            return name
        n = self._counter.get(hint, 0)
        self._counter[hint] = n + 1
        return "%s%d" % (hint, n)

    @classmethod
    def _get_current(cls):
        if not hasattr(NameManager._current, "value"):
            NameManager._current.value = NameManager()
        return NameManager._current.value

    def __enter__(self):
        self._old_manager = NameManager._get_current()
        NameManager._current.value = self
        return self

    def __exit__(self, ptype, value, trace):
        assert self._old_manager is not None
        NameManager._current.value = self._old_manager


class Prefix(NameManager):
    """Every name (given or generated) created in the scope gets ``prefix`` prepended."""

    def __init__(self, prefix):
        super(Prefix, self).__init__()
        self._prefix = prefix

    def get(self, name, hint):
        return self._prefix + super(Prefix, self).get(name, hint)
