"""Learning-rate schedulers (reference: python/mxnet/lr_scheduler.py:41-281) plus the dynamic-mini-batch
rule of the paper behind the fork (Lin et al., "Dynamic Mini-batch SGD for Elastic Distributed Training",
arXiv 1904.12043), which the reference tree itself never implemented.

Every scheduler is a callable ``sched(num_update) -> lr``; ``Optimizer.__init__`` overwrites
``sched.base_lr`` with its ``learning_rate`` (optimizer.py:92-93).

Provenance: adapted from Apache MXNet (incubating) 1.3.1 ``python/mxnet/lr_scheduler.py`` (Apache License 2.0).  The public API, argument
names and control flow follow the original so that user scripts keep working; the tensor maths runs on torch /
the mxb200 kernels.  This file is API-compatibility glue, not original B200 work ("Provenance").
"""
import logging
import math

__all__ = ["LRScheduler", "FactorScheduler", "MultiFactorScheduler", "PolyScheduler", "CosineScheduler",
           "WarmupScheduler", "DynamicMiniBatchScheduler"]


class LRScheduler(object):
    def __init__(self, base_lr=0.01, warmup_steps=0, warmup_begin_lr=0, warmup_mode="linear"):
        self.base_lr = base_lr
        assert isinstance(warmup_steps, int)
        self.warmup_steps = warmup_steps
        self.warmup_final_lr = base_lr
        self.warmup_begin_lr = warmup_begin_lr
        if self.warmup_begin_lr > self.warmup_final_lr:
            raise ValueError("Base lr has to be higher than warmup_begin_lr")
        if self.warmup_steps < 0:
            raise ValueError("Warmup steps has to be positive or 0")
        if warmup_mode not in ["linear", "constant"]:
            raise ValueError("Supports only linear and constant modes of warmup")
        self.warmup_mode = warmup_mode

    def get_warmup_lr(self, num_update):
        assert num_update < self.warmup_steps
        if self.warmup_mode == "linear":
            increase = (self.warmup_final_lr - self.warmup_begin_lr) * float(num_update) / float(self.warmup_steps)
            return self.warmup_begin_lr + increase
        return self.warmup_begin_lr

    def __call__(self, num_update):
        raise NotImplementedError("must override this")


class FactorScheduler(LRScheduler):
    """lr *= factor every ``step`` updates, floored at stop_factor_lr."""

    def __init__(self, step, factor=1, stop_factor_lr=1e-8, base_lr=0.01, warmup_steps=0, warmup_begin_lr=0,
                 warmup_mode="linear"):
        super(FactorScheduler, self).__init__(base_lr, warmup_steps, warmup_begin_lr, warmup_mode)
        if step < 1:
            raise ValueError("Schedule step must be greater or equal than 1 round")
        if factor > 1.0:
            raise ValueError("Factor must be no more than 1 to make lr reduce")
        self.step, self.factor, self.stop_factor_lr = step, factor, stop_factor_lr
        self.count = 0

    def __call__(self, num_update):
        if num_update < self.warmup_steps:
            return self.get_warmup_lr(num_update)
        while num_update > self.count + self.step:
            self.count += self.step
            self.base_lr *= self.factor
            if self.base_lr < self.stop_factor_lr:
                self.base_lr = self.stop_factor_lr
                logging.info("Update[%d]: now learning rate arrived at %0.5e, will not change in the future",
                             num_update, self.base_lr)
            else:
                logging.info("Update[%d]: Change learning rate to %0.5e", num_update, self.base_lr)
        return self.base_lr


class MultiFactorScheduler(LRScheduler):
    """lr *= factor at each update count listed in ``step``."""

    def __init__(self, step, factor=1, base_lr=0.01, warmup_steps=0, warmup_begin_lr=0, warmup_mode="linear"):
        super(MultiFactorScheduler, self).__init__(base_lr, warmup_steps, warmup_begin_lr, warmup_mode)
        assert isinstance(step, list) and len(step) >= 1
        for i, _step in enumerate(step):
            if i != 0 and step[i] <= step[i - 1]:
                raise ValueError("Schedule step must be an increasing integer list")
            if _step < 1:
                raise ValueError("Schedule step must be greater or equal than 1 round")
        if factor > 1.0:
            raise ValueError("Factor must be no more than 1 to make lr reduce")
        self.step, self.cur_step_ind, self.factor, self.count = step, 0, factor, 0

    def __call__(self, num_update):
        if num_update < self.warmup_steps:
            return self.get_warmup_lr(num_update)
        while self.cur_step_ind <= len(self.step) - 1:
            if num_update > self.step[self.cur_step_ind]:
                self.count = self.step[self.cur_step_ind]
                self.cur_step_ind += 1
                self.base_lr *= self.factor
                logging.info("Update[%d]: Change learning rate to %0.5e", num_update, self.base_lr)
            else:
                return self.base_lr
        return self.base_lr


class PolyScheduler(LRScheduler):
    """final + (base-final) * (1 - t/T)^pwr after warm-up."""

    def __init__(self, max_update, base_lr=0.01, pwr=2, final_lr=0, warmup_steps=0, warmup_begin_lr=0,
                 warmup_mode="linear"):
        super(PolyScheduler, self).__init__(base_lr, warmup_steps, warmup_begin_lr, warmup_mode)
        assert isinstance(max_update, int)
        if max_update < 1:
            raise ValueError("maximum number of updates must be strictly positive")
        self.power, self.base_lr_orig, self.max_update, self.final_lr = pwr, self.base_lr, max_update, final_lr
        self.max_steps = self.max_update - self.warmup_steps

    def __call__(self, num_update):
        if num_update < self.warmup_steps:
            return self.get_warmup_lr(num_update)
        if num_update <= self.max_update:
            self.base_lr = self.final_lr + (self.base_lr_orig - self.final_lr) * \
                pow(1 - float(num_update - self.warmup_steps) / float(self.max_steps), self.power)
        return self.base_lr


class CosineScheduler(LRScheduler):
    def __init__(self, max_update, base_lr=0.01, final_lr=0, warmup_steps=0, warmup_begin_lr=0, warmup_mode="linear"):
        super(CosineScheduler, self).__init__(base_lr, warmup_steps, warmup_begin_lr, warmup_mode)
        assert isinstance(max_update, int)
        if max_update < 1:
            raise ValueError("maximum number of updates must be strictly positive")
        self.base_lr_orig, self.max_update, self.final_lr = base_lr, max_update, final_lr
        self.max_steps = self.max_update - self.warmup_steps

    def __call__(self, num_update):
        if num_update < self.warmup_steps:
            return self.get_warmup_lr(num_update)
        if num_update <= self.max_update:
            self.base_lr = self.final_lr + (self.base_lr_orig - self.final_lr) * \
                (1 + math.cos(math.pi * (num_update - self.warmup_steps) / self.max_steps)) / 2
        return self.base_lr


class WarmupScheduler(LRScheduler):
    """Linear warm-up wrapped around any other scheduler (the example re-implements this,
    example/dynamic-training/train_resnet.py:22-229)."""

    def __init__(self, scheduler, warmup_steps, warmup_begin_lr=0.0):
        super(WarmupScheduler, self).__init__(getattr(scheduler, "base_lr", 0.01), 0, 0)
        self.scheduler, self.wsteps, self.wbegin = scheduler, int(warmup_steps), warmup_begin_lr

    def __call__(self, num_update):
        self.scheduler.base_lr = getattr(self.scheduler, "base_lr", self.base_lr)
        if num_update < self.wsteps:
            target = self.scheduler(self.wsteps)
            return self.wbegin + (target - self.wbegin) * num_update / max(1, self.wsteps)
        return self.scheduler(num_update)


class DynamicMiniBatchScheduler(LRScheduler):
    """Dynamic mini-batch SGD learning-rate rule for elastic training.

    The reference keeps the GLOBAL batch constant when workers come and go (per-worker batch =
    global/num_workers, example/dynamic-training/train_resnet.py:315-317), so no lr change is needed; that
    is ``mode='constant_global_batch'`` (factor 1).  With a constant PER-WORKER batch the global batch B
    changes with the membership and the paper's rule applies: lr follows linear scaling lr0*B/B0, and to
    avoid the noise-scale jump at a membership change the effective lr is moved to the new target smoothly
    over ``smooth_steps`` updates (momentum compensation: with heavy-ball momentum mu the effective step is
    lr/(1-mu), so an abrupt k-fold lr change is a k-fold jump of an already accumulated velocity).
    """

    def __init__(self, scheduler, base_batch, mode="linear", smooth_steps=0, momentum=0.0, per_worker_batch=None):
        super(DynamicMiniBatchScheduler, self).__init__(getattr(scheduler, "base_lr", 0.01), 0, 0)
        assert mode in ("linear", "sqrt", "constant_global_batch")
        self.scheduler, self.base_batch, self.mode = scheduler, float(base_batch), mode
        # constant PER-WORKER batch: the frameworks' membership hooks (gluon.Trainer._on_membership_change,
        # BaseModule._on_membership_change) compute the new global batch as per_worker_batch * num_workers
        self.per_worker_batch = per_worker_batch
        self.smooth_steps, self.momentum = int(smooth_steps), momentum
        self.cur_factor = 1.0
        self._from, self._to, self._t0 = 1.0, 1.0, 0
        self._last_update = 0

    def _target(self, batch):
        r = float(batch) / self.base_batch
        return {"linear": r, "sqrt": math.sqrt(r), "constant_global_batch": 1.0}[self.mode]

    def on_membership_change(self, new_global_batch, num_update=None):
        """Call at the epoch boundary where the worker set (hence the global batch) changed."""
        self._from = self.cur_factor
        self._to = self._target(new_global_batch)
        self._t0 = self._last_update if num_update is None else num_update
        if self.smooth_steps == 0:
            self.cur_factor = self._to
        return self._to

    def momentum_correction(self):
        """Factor to multiply the momentum buffer with at an abrupt change (Goyal et al. 2017, remark 2)."""
        return self._to / self._from if self._from else 1.0

    def __call__(self, num_update):
        self._last_update = num_update
        if hasattr(self.scheduler, "base_lr") and not getattr(self, "_synced", False):
            self.scheduler.base_lr = self.base_lr
            if hasattr(self.scheduler, "base_lr_orig"):
                self.scheduler.base_lr_orig = self.base_lr
            self._synced = True
        if self.smooth_steps > 0 and self.cur_factor != self._to:
            a = min(1.0, max(0.0, (num_update - self._t0) / float(self.smooth_steps)))
            self.cur_factor = self._from + (self._to - self._from) * a
        return self.scheduler(num_update) * self.cur_factor
