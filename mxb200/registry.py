"""Generic class registries (reference: python/mxnet/registry.py): ``get_register_func`` / ``get_alias_func`` /
``get_create_func`` build the register / alias / create trio used by optimizers, initializers, metrics.

Provenance: adapted from Apache MXNet (incubating) 1.3.1 ``python/mxnet/registry.py`` (Apache License 2.0).  The public API, argument
names and control flow follow the original so that user scripts keep working; the tensor maths runs on torch /
the mxb200 kernels.  This file is API-compatibility glue, not original B200 work ("Provenance").
"""
import json
import warnings

_REGISTRY = {}


def get_registry(base_class):
    if base_class not in _REGISTRY:
        _REGISTRY[base_class] = {}
    return _REGISTRY[base_class].copy()


def get_register_func(base_class, nickname):
    if base_class not in _REGISTRY:
        _REGISTRY[base_class] = {}
    registry = _REGISTRY[base_class]

    def register(klass, name=None):
        assert issubclass(klass, base_class), "Can only register subclass of %s" % base_class.__name__
        name = (name or klass.__name__).lower()
        if name in registry:
            warnings.warn("\033[91mNew %s %s.%s registered with name %s is overriding existing %s %s.%s\033[0m" % (
                nickname, klass.__module__, klass.__name__, name, nickname, registry[name].__module__, registry[name].__name__),
                UserWarning, stacklevel=2)
        registry[name] = klass
        return klass
    register.__doc__ = "Register %s to the %s factory" % (nickname, nickname)
    return register


def get_alias_func(base_class, nickname):
    register = get_register_func(base_class, nickname)

    def alias(*aliases):
        def reg(klass):
            for name in aliases:
                register(klass, name)
            return klass
        return reg
    return alias


def get_create_func(base_class, nickname):
    if base_class not in _REGISTRY:
        _REGISTRY[base_class] = {}
    registry = _REGISTRY[base_class]

    def create(*args, **kwargs):
        if len(args):
            name, args = args[0], args[1:]
        else:
            name = kwargs.pop(nickname)
        if isinstance(name, base_class):
            assert len(args) == 0 and len(kwargs) == 0, "%s is already an instance. Additional arguments are invalid" % nickname
            return name
        if isinstance(name, dict):
            return create(**name)
        assert isinstance(name, str), "%s must be of string type" % nickname
        if name.startswith("["):
            assert not args and not kwargs
            name, kwargs = json.loads(name)
            return create(name, **kwargs)
        if name.startswith("{"):
            assert not args and not kwargs
            return create(**json.loads(name))
        name = name.lower()
        assert name in registry, "%s is not registered. Please register with %s.register first" % (str(name), nickname)
        return registry[name](*args, **kwargs)
    create.__doc__ = "Create a %s instance from config." % nickname
    return create
