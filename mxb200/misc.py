"""Deprecated learning-rate helpers kept for old scripts (reference: python/mxnet/misc.py); use mx.lr_scheduler.

Provenance: adapted from Apache MXNet (incubating) 1.3.1 ``python/mxnet/misc.py`` (Apache License 2.0).  The public API, argument
names and control flow follow the original so that user scripts keep working; the tensor maths runs on torch /
the mxb200 kernels.  This file is API-compatibility glue, not original B200 work ("Provenance").
"""
import logging
import math


class LearningRateScheduler(object):
    def __init__(self):
        self.base_lr = 0.01

    def __call__(self, iteration):
        raise NotImplementedError("must override this")


class FactorScheduler(LearningRateScheduler):
    """lr = base_lr * factor^(floor(iteration / step))."""

    def __init__(self, step, factor=0.1):
        super(FactorScheduler, self).__init__()
        if step < 1:
            raise ValueError("Schedule step must be greater or equal than 1 round")
        if factor >= 1.0:
            raise ValueError("Factor must be less than 1 to make lr reduce")
        self.step, self.factor = step, factor
        self.old_lr = self.base_lr
        self.init = False

    def __call__(self, iteration):
        if not self.init:
            self.init = True
            self.old_lr = self.base_lr
        lr = self.base_lr * math.pow(self.factor, int(iteration / self.step))
        if lr != self.old_lr:
            self.old_lr = lr
            logging.info("At Iteration [%d]: Swith to new learning rate %.5f", iteration, lr)
        return lr
