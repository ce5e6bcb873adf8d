"""Contrib namespace (reference: python/mxnet/contrib/): text, svrg_optimization, io, quantization, tensorboard and the
``ndarray`` / ``symbol`` / ``autograd`` aliases.  ``onnx`` and ``tensorrt`` need external runtimes that are not part of
this image; importing them raises a clear error."""
from . import autograd  # noqa: F401  (the pre-1.0 interface: set_is_training / train_section / grad_and_loss)
from ..ndarray import contrib as ndarray  # noqa: F401
nd = ndarray
from . import io, quantization, svrg_optimization, tensorboard, text  # noqa: F401


def __getattr__(name):
    if name in ("onnx", "tensorrt"):
        raise ImportError("mxb200.contrib.%s needs the %s runtime, which is not available in this build" % (name, name))
    if name == "symbol":
        from ..symbol import contrib as _sc
        return _sc
    raise AttributeError(name)
