"""Helper optimizers of SVRG training with a KVStore (reference: python/mxnet/contrib/svrg_optimization/svrg_optimizer.py).

With several workers the full gradients are accumulated in the store under keys whose names contain ``full``; those
keys must be *assigned* (the pushed value replaces the stored one) while every ordinary parameter key is updated by the
user's optimizer.  ``_SVRGOptimizer`` dispatches between the two by key name."""
from ... import optimizer as _opt

__all__ = ["_AssignmentOptimizer", "_SVRGOptimizer"]


@_opt.register
class _AssignmentOptimizer(_opt.Optimizer):
    """weight <- grad: the stored value becomes whatever was pushed."""

    def create_state(self, index, weight):
        return None

    def update(self, index, weight, grad, state):
        weight[:] = grad


@_opt.register
class _SVRGOptimizer(_opt.Optimizer):
    """``default_optimizer`` (name or instance) for ordinary keys, assignment for the full-gradient keys."""

    _BASE_PARAMS = ("rescale_grad", "param_idx2name", "wd", "clip_gradient", "learning_rate", "lr_scheduler", "sym",
                    "begin_num_update", "multi_precision", "param_dict")

    def __init__(self, default_optimizer, **kwargs):
        super(_SVRGOptimizer, self).__init__(**self._check_params(**kwargs))
        self.default_opt = _opt.create(default_optimizer, **kwargs) if isinstance(default_optimizer, str) else default_optimizer
        self.aux_opt = _opt.create(_AssignmentOptimizer.__name__)

    @classmethod
    def _check_params(cls, **kwargs):
        """The subset of kwargs the Optimizer base class understands."""
        return {k: v for k, v in kwargs.items() if k in cls._BASE_PARAMS}

    def _check_index(self, index):
        return index if index in self.idx2name.values() else self.idx2name[index]

    def _pick(self, index):
        return self.aux_opt if "full" in self._check_index(index) else self.default_opt

    def create_state(self, index, weight):
        return self._pick(index).create_state(index, weight)

    def update(self, index, weight, grad, state):
        self._pick(index).update(index, weight, grad, state)
