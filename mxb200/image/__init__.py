"""Image IO and augmentation (reference: python/mxnet/image/image.py; C++ src/io/image_aug_default.cc)."""
from .image import *  # noqa: F401,F403
