"""contrib.text: vocabulary and token-embedding utilities (reference: python/mxnet/contrib/text/)."""
from . import embedding, utils, vocab  # noqa: F401
from .vocab import Vocabulary  # noqa: F401
