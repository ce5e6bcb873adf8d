"""contrib.svrg_optimization: Stochastic Variance Reduced Gradient training (reference:
python/mxnet/contrib/svrg_optimization/{svrg_module,svrg_optimizer}.py; Johnson & Zhang 2013)."""
from .svrg_module import SVRGModule  # noqa: F401
from . import svrg_optimizer  # noqa: F401,E402
