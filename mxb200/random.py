"""mx.random (reference: python/mxnet/random.py)."""
from .ndarray.random import seed, uniform, normal, randn, randint, exponential, gamma, poisson, multinomial, shuffle  # noqa: F401
