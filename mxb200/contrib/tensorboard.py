"""contrib.tensorboard.LogMetricsCallback (reference: python/mxnet/contrib/tensorboard.py).  Writes scalars through
torch.utils.tensorboard when available, otherwise to a plain ``scalars.tsv`` in the logging directory."""
import os
import time


class LogMetricsCallback(object):
    def __init__(self, logging_dir, prefix=None):
        self.prefix = prefix
        self.step = 0
        os.makedirs(logging_dir, exist_ok=True)
        try:
            from torch.utils.tensorboard import SummaryWriter
            self._writer = SummaryWriter(logging_dir)
            self._tsv = None
        except Exception:  # noqa: BLE001
            self._writer = None
            self._tsv = open(os.path.join(logging_dir, "scalars.tsv"), "a")

    def __call__(self, param):
        if param.eval_metric is None:
            return
        self.step += 1
        for name, value in param.eval_metric.get_name_value():
            if self.prefix is not None:
                name = "%s-%s" % (self.prefix, name)
            if self._writer is not None:
                self._writer.add_scalar(name, value, self.step)
            else:
                self._tsv.write("%f\t%d\t%s\t%r\n" % (time.time(), self.step, name, value))
                self._tsv.flush()
