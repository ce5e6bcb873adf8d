"""``mx.sym`` namespace: every op registered in ``mx.nd`` is available symbolically."""
import sys as _sys
import types as _types

from .symbol import Symbol, var, Variable, Group, load, load_json, zeros, ones, _create  # noqa: F401
from .symbol import full, arange, eye, pow, power, maximum, minimum, hypot  # noqa: F401,A004
from ..ndarray import op as _ndop


def _lookup_op(name):
    if name == "Custom" and _ndop.get(name) is None:
        import importlib
        importlib.import_module(__name__.rsplit(".", 1)[0] + ".operator")      # registers the Custom operator
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


# ----------------------------------------------------------------------------- sub-namespaces (python/mxnet/symbol/{random,linalg,sparse,image,_internal}.py)
class _Prefixed(_types.ModuleType):
    """mx.sym.<ns>.<op> -> operator ``<prefix><op>`` (falls back to the plain name, e.g. mx.sym.sparse.dot)."""

    def __init__(self, name, prefixes, fallback=True):
        super(_Prefixed, self).__init__(name)
        self._prefixes, self._fallback = prefixes, fallback

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        for pre in self._prefixes:
            f = _lookup_op(pre + name)
            if f is not None:
                return f
        f = _lookup_op(name) if self._fallback else None
        if f is None:
            raise AttributeError("module %r has no attribute %r" % (self.__name__, name))
        return f


linalg = _Prefixed(__name__ + ".linalg", ("_linalg_", "linalg_"), fallback=False)
sparse = _Prefixed(__name__ + ".sparse", ("_sparse_",))
image = _Prefixed(__name__ + ".image", ("_image_",), fallback=False)
_internal = _Prefixed(__name__ + "._internal", ("",))
op = _Prefixed(__name__ + ".op", ("",))
for _m in (linalg, sparse, image, _internal, op):
    _sys.modules[_m.__name__] = _m


class _SymRandom(_types.ModuleType):
    """mx.sym.random: scalar parameters -> ``_random_*`` nodes, Symbol parameters -> ``_sample_*`` nodes
    (python/mxnet/symbol/random.py:29-45)."""

    @staticmethod
    def _helper(name, params, names, shape, dtype, kwargs):
        if shape is not None:
            kwargs["shape"] = shape
        if dtype is not None:
            kwargs["dtype"] = dtype
        if all(isinstance(p, Symbol) for p in params):
            return _create("_sample_" + name, list(params), kwargs)
        if any(isinstance(p, Symbol) for p in params):
            raise ValueError("Distribution parameters must all have the same type, but got both %s" % str([type(p) for p in params]))
        return _create("_random_" + name, [], dict(zip(names, params), **kwargs))

    def uniform(self, low=0, high=1, shape=None, dtype=None, **kwargs):
        return self._helper("uniform", [low, high], ["low", "high"], shape, dtype, kwargs)

    def normal(self, loc=0, scale=1, shape=None, dtype=None, **kwargs):
        return self._helper("normal", [loc, scale], ["loc", "scale"], shape, dtype, kwargs)

    def poisson(self, lam=1, shape=None, dtype=None, **kwargs):
        return self._helper("poisson", [lam], ["lam"], shape, dtype, kwargs)

    def exponential(self, scale=1, shape=None, dtype=None, **kwargs):
        return self._helper("exponential", [1.0 / scale], ["lam"], shape, dtype, kwargs)

    def gamma(self, alpha=1, beta=1, shape=None, dtype=None, **kwargs):
        return self._helper("gamma", [alpha, beta], ["alpha", "beta"], shape, dtype, kwargs)

    def negative_binomial(self, k=1, p=1, shape=None, dtype=None, **kwargs):
        return self._helper("negative_binomial", [k, p], ["k", "p"], shape, dtype, kwargs)

    def generalized_negative_binomial(self, mu=1, alpha=1, shape=None, dtype=None, **kwargs):
        return self._helper("generalized_negative_binomial", [mu, alpha], ["mu", "alpha"], shape, dtype, kwargs)

    def multinomial(self, data, shape=None, get_prob=True, dtype="int32", **kwargs):
        if shape is not None:
            kwargs["shape"] = shape
        return _create("_sample_multinomial", [data], dict(get_prob=get_prob, dtype=dtype, **kwargs))

    def shuffle(self, data, **kwargs):
        return _create("_shuffle", [data], kwargs)


random = _SymRandom(__name__ + ".random")
_sys.modules[random.__name__] = random


def _sym_rand_zipfian(true_classes, num_sampled, range_max):
    """Symbolic log-uniform candidate sampler (symbol/contrib.py:rand_zipfian of the reference); see nd.contrib.rand_zipfian."""
    import math as _math
    log_range = _math.log(range_max + 1)
    r = random.uniform(0, log_range, shape=(num_sampled,), dtype="float64")
    sampled = _create("_mod_scalar", [_create("Cast", [_create("exp", [r], {}) - 1], {"dtype": "int64"})], {"scalar": range_max})

    def expected(cls):
        c = _create("Cast", [cls], {"dtype": "float64"})
        return _create("log", [(c + 2.0) / (c + 1.0)], {}) / log_range * num_sampled
    return sampled, expected(true_classes), expected(sampled)


contrib.rand_zipfian = _sym_rand_zipfian
