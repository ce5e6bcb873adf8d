"""Contrib optimizers (reference: python/mxnet/optimizer/contrib.py)."""
from ..ndarray import zeros, op as ndop
from .optimizer import Optimizer, register

__all__ = ["GroupAdaGrad"]


@register
class GroupAdaGrad(Optimizer):
    """AdaGrad with ONE accumulator per row (embedding tables): history[r] += mean(g[r]^2);
    w[r] -= lr * g[r] / sqrt(history[r] + eps).  Weight decay is not supported (contrib.py:30-100 of the reference)."""

    def __init__(self, eps=1e-5, **kwargs):
        super(GroupAdaGrad, self).__init__(**kwargs)
        self.float_stable_eps = eps

    def create_state(self, index, weight):
        assert len(weight.shape) == 2
        return zeros((weight.shape[0], 1), ctx=weight.context, dtype=weight.dtype)

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        assert wd == 0, "Weight decay is not supported for GroupAdaGrad"
        ndop.get("_contrib_group_adagrad_update")(weight, grad, state, lr=lr, rescale_grad=self.rescale_grad,
                                                  clip_gradient=-1.0 if self.clip_gradient is None else self.clip_gradient,
                                                  epsilon=self.float_stable_eps, out=weight)
