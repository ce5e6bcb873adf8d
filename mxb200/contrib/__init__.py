"""Contrib namespace (reference: python/mxnet/contrib/): autograd alias, quantization / text / onnx are out of the
single-box product's scope and raise a clear error when used."""
from .. import autograd  # noqa: F401
from ..ndarray import contrib as ndarray  # noqa: F401
