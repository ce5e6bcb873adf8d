"""Vision model zoo (reference: python/mxnet/gluon/model_zoo/vision/__init__.py)."""
from .resnet import *  # noqa: F401,F403
from .others import *  # noqa: F401,F403
from . import resnet as _resnet, others as _others


def get_model(name, **kwargs):
    models = {}
    for mod in (_resnet, _others):
        for n in mod.__all__:
            f = getattr(mod, n)
            if callable(f) and n[0].islower():
                models[n] = f
    name = name.lower()
    if name not in models:
        raise ValueError("Model %s is not supported. Available options are\n\t%s" % (name, "\n\t".join(sorted(models.keys()))))
    return models[name](**kwargs)
