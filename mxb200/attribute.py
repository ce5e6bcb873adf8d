"""Attribute scope (reference: python/mxnet/attribute.py)."""
import threading


class _ScopeMeta(type):
    """``AttrScope.current``: the calling thread's innermost scope, readable and assignable on the class
    (python/mxnet/attribute.py of the reference: classproperty over a thread-local)."""

    @property
    def current(cls):
        cur = getattr(cls._current, "value", None)
        if cur is None:
            cur = cls._current.value = cls()
        return cur

    @current.setter
    def current(cls, value):
        cls._current.value = value


class AttrScope(object, metaclass=_ScopeMeta):
    _current = threading.local()

    def __init__(self, **kwargs):
        self._old_scope = None
        for value in kwargs.values():
            if not isinstance(value, str):
                raise ValueError("Attributes need to be string")
        self._attr = kwargs

    def get(self, attr):
        if self._attr:
            ret = self._attr.copy()
            if attr:
                ret.update(attr)
            return ret
        return attr if attr else {}

    def __enter__(self):
        self._old_scope = getattr(AttrScope._current, "value", None)
        attr = (self._old_scope._attr.copy() if self._old_scope else {})
        attr.update(self._attr)
        self._attr = attr
        AttrScope._current.value = self
        return self

    def __exit__(self, *a):
        AttrScope._current.value = self._old_scope


def current_attrs():
    """Attributes of the innermost active ``AttrScope`` (empty dict outside any scope)."""
    sc = getattr(AttrScope._current, "value", None)
    return dict(sc._attr) if sc is not None else {}
