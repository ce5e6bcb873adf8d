"""Training callbacks (reference: python/mxnet/callback.py).

Provenance: adapted from Apache MXNet (incubating) 1.3.1 ``python/mxnet/callback.py`` (Apache License 2.0).  The public API, argument
names and control flow follow the original so that user scripts keep working; the tensor maths runs on torch /
the mxb200 kernels.  This file is API-compatibility glue, not original B200 work ("Provenance").
"""
import logging
import math
import time

from .model import save_checkpoint
from .base import getenv

__all__ = ["module_checkpoint", "do_checkpoint", "log_train_metric", "Speedometer", "ProgressBar", "LogValidationMetricsCallback"]


def module_checkpoint(mod, prefix, period=1, save_optimizer_states=False):
    period = int(max(1, period))

    def _callback(iter_no, sym=None, arg=None, aux=None):
        if (iter_no + 1) % period == 0:
            mod.save_checkpoint(prefix, iter_no + 1, save_optimizer_states)
    return _callback


def do_checkpoint(prefix, period=1):
    """Epoch-end callback saving prefix-symbol.json / prefix-%04d.params every ``period`` epochs (callback.py:55-90)."""
    period = int(max(1, period))

    def _callback(iter_no, sym, arg, aux):
        if (iter_no + 1) % period == 0:
            # written by the dependency engine while the next epoch trains (fit() / nd.load / waitall wait for it)
            save_checkpoint(prefix, iter_no + 1, sym, arg, aux, async_write=bool(getenv("MXB200_ASYNC_CHECKPOINT", True)))
    return _callback


def log_train_metric(period, auto_reset=False):
    def _callback(param):
        if param.nbatch % period == 0 and param.eval_metric is not None:
            name_value = param.eval_metric.get_name_value()
            for name, value in name_value:
                logging.info("Iter[%d] Batch[%d] Train-%s=%f", param.epoch, param.nbatch, name, value)
            if auto_reset:
                param.eval_metric.reset()
    return _callback


class Speedometer(object):
    """Logs samples/sec every ``frequent`` batches: frequent*batch_size / wall-time (callback.py:120-177)."""

    def __init__(self, batch_size, frequent=50, auto_reset=True):
        self.batch_size = batch_size
        self.frequent = frequent
        self.init = False
        self.tic = 0
        self.last_count = 0
        self.auto_reset = auto_reset
        self.last_speed = None

    def _batch_size(self):
        return self.batch_size

    def __call__(self, param):
        count = param.nbatch
        if self.last_count > count:
            self.init = False
        self.last_count = count
        if self.init:
            if count % self.frequent == 0:
                speed = self.frequent * self._batch_size() / (time.time() - self.tic)
                self.last_speed = speed
                if param.eval_metric is not None:
                    name_value = param.eval_metric.get_name_value()
                    if self.auto_reset:
                        param.eval_metric.reset()
                    msg = "Epoch[%d] Batch [%d]\tSpeed: %.2f samples/sec"
                    msg += "\t%s=%f" * len(name_value)
                    logging.info(msg, param.epoch, count, speed, *sum(name_value, ()))
                else:
                    logging.info("Iter[%d] Batch [%d]\tSpeed: %.2f samples/sec", param.epoch, count, speed)
                self.tic = time.time()
        else:
            self.init = True
            self.tic = time.time()


class ElasticSpeedometer(Speedometer):
    """Speedometer whose local batch follows the live worker count: global_batch / kv.num_workers
    (example/dynamic-training/train_resnet.py:381-390)."""

    def __init__(self, kv, global_batch_size, frequent=50, auto_reset=True):
        super(ElasticSpeedometer, self).__init__(global_batch_size, frequent, auto_reset)
        self.kv = kv
        self.global_batch_size = global_batch_size

    def _batch_size(self):
        return self.global_batch_size // max(1, self.kv.num_workers)


class ProgressBar(object):
    def __init__(self, total, length=80):
        self.bar_len = length
        self.total = total

    def __call__(self, param):
        count = param.nbatch
        filled_len = int(round(self.bar_len * count / float(self.total)))
        percents = math.ceil(100.0 * count / float(self.total))
        prog_bar = "=" * filled_len + "-" * (self.bar_len - filled_len)
        logging.info("[%s] %s%s\r", prog_bar, percents, "%")


class LogValidationMetricsCallback(object):
    def __call__(self, param):
        if not param.eval_metric:
            return
        name_value = param.eval_metric.get_name_value()
        for name, value in name_value:
            logging.info("Epoch[%d] Validation-%s=%f", param.epoch, name, value)
