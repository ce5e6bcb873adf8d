"""Library information (reference: python/mxnet/libinfo.py)."""
import os

__version__ = "0.1.0"


def find_lib_path():
    here = os.path.dirname(os.path.abspath(__file__))
    return [os.path.join(here, n) for n in ("libmxb200_host.so", "libmxb200_cuda.so") if os.path.exists(os.path.join(here, n))]
