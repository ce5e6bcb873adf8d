"""dtype codes shared with csrc/cuda/common.cuh (mshadow type flags; bf16 = 12)."""
import torch

_CODE = {torch.float32: 0, torch.float16: 2, torch.bfloat16: 12, torch.int64: 6, torch.int32: 4}


def code(dtype):
    try:
        return _CODE[dtype]
    except KeyError:
        raise TypeError("dtype %s is not supported by the native kernels" % dtype)
