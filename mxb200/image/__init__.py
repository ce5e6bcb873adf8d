"""Image IO and augmentation (reference: python/mxnet/image/{image,detection}.py; C++ src/io/image_aug_default.cc)."""
from .image import *  # noqa: F401,F403
from . import detection  # noqa: F401
from .detection import *  # noqa: F401,F403
