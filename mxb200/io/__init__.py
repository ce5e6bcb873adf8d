"""Data iterators (reference: python/mxnet/io/)."""
from .io import *  # noqa: F401,F403
from .io import DataDesc, DataBatch, DataIter  # noqa: F401
from .device_prefetch import DevicePrefetcher  # noqa: F401,E402
from . import gpu_pipeline  # noqa: F401,E402
from .gpu_pipeline import GPUImageRecordIter  # noqa: F401,E402
