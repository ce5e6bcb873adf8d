"""The pre-1.0 autograd interface (reference: python/mxnet/contrib/autograd.py), kept for old scripts; new code uses
``mx.autograd``."""
import functools

from .. import autograd as _ag
from ..ndarray import NDArray

__all__ = ["set_is_training", "TrainingStateScope", "train_section", "test_section", "mark_variables", "backward",
           "compute_gradient", "grad_and_loss", "grad"]


def set_is_training(is_train):
    prev = _ag.is_training()
    _ag.set_training(is_train)
    _ag.set_recording(is_train)
    return prev


class TrainingStateScope(object):
    def __init__(self, enter_state):
        self._enter_state = enter_state
        self._prev = None

    def __enter__(self):
        self._prev = set_is_training(self._enter_state)

    def __exit__(self, ptype, value, trace):
        if self._prev != self._enter_state:
            set_is_training(self._prev)


def train_section():
    return TrainingStateScope(True)


def test_section():
    return TrainingStateScope(False)


mark_variables = _ag.mark_variables


def backward(outputs, out_grads=None, retain_graph=False):
    _ag.backward(outputs, out_grads, retain_graph)


compute_gradient = backward


def grad_and_loss(func, argnum=None):
    """Returns a function computing (gradients w.r.t. the selected positional arguments, outputs of func)."""
    @functools.wraps(func)
    def wrapped(*args):
        variables = args
        if argnum is not None:
            idx = [argnum] if isinstance(argnum, int) else list(argnum)
            variables = [args[i] for i in idx]
        for x in variables:
            assert isinstance(x, NDArray), "type of autograd input should be NDArray."
        grads = [x.zeros_like() for x in variables]
        mark_variables(variables, grads)
        with train_section():
            outputs = func(*args)
        backward([outputs] if isinstance(outputs, NDArray) else outputs)
        return grads, outputs
    return wrapped


def grad(func, argnum=None):
    g = grad_and_loss(func, argnum)

    @functools.wraps(g)
    def wrapped(*args):
        return g(*args)[0]
    return wrapped
