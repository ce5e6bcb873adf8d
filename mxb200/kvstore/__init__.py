from .kvstore import KVStore, KVStoreLocal, create, MAX_ALLOWED_KEY_FOR_UPDATE  # noqa: F401
from .compression import GradientCompression  # noqa: F401
