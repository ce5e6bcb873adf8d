"""Module API (reference: python/mxnet/module/)."""
from .base_module import BaseModule  # noqa: F401
from .module import Module  # noqa: F401
from .base_data_iterator import BaseDataIterator, ElasticDataIterator  # noqa: F401
from .extra_modules import BucketingModule, SequentialModule, PythonModule, PythonLossModule  # noqa: F401
