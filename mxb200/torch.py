"""Bridge to PyTorch (the reference's python/mxnet/torch.py wraps Lua-Torch functions through a plugin; this framework
runs ON PyTorch, so the bridge is the zero-copy conversion pair)."""
from .ndarray.ndarray import NDArray, from_torch  # noqa: F401


def to_torch(arr):
    """The torch.Tensor that backs ``arr`` (shares memory)."""
    return arr.as_torch() if isinstance(arr, NDArray) else arr
