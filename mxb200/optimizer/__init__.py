from .optimizer import *  # noqa: F401,F403
from . import optimizer  # noqa: F401
from .optimizer import Optimizer, Updater, get_updater, create, register  # noqa: F401
from . import contrib  # noqa: F401,E402
