"""Training callbacks that collect metrics into pandas DataFrames for notebooks (reference:
python/mxnet/notebook/callback.py: PandasLogger, LiveLearningCurve; the bokeh live plots need a notebook front-end and
degrade to data collection here)."""
import datetime
import time

try:
    import pandas as pd
except ImportError:  # pragma: no cover
    pd = None


def _add_new_columns(df, new_columns):
    for c in new_columns:
        if c not in df.columns:
            df[c] = [None] * len(df)
    return df


class PandasLogger(object):
    """Logs train / eval / epoch metrics into three DataFrames: use ``callback_args()`` as ``**kwargs`` of ``fit``."""

    def __init__(self, batch_size, frequent=50):
        if pd is None:
            raise ImportError("PandasLogger needs pandas")
        self.batch_size, self.frequent = batch_size, frequent
        self._dataframes = {"train": pd.DataFrame(), "eval": pd.DataFrame(), "epoch": pd.DataFrame()}
        self.last_time = time.time()
        self.start_time = datetime.datetime.now()
        self.last_epoch_time = datetime.datetime.now()

    @property
    def train_df(self):
        return self._dataframes["train"]

    @property
    def eval_df(self):
        return self._dataframes["eval"]

    @property
    def epoch_df(self):
        return self._dataframes["epoch"]

    @property
    def all_dataframes(self):
        return self._dataframes

    def elapsed(self):
        return datetime.datetime.now() - self.start_time

    def append_metrics(self, metrics, df_name):
        df = self._dataframes[df_name]
        df = _add_new_columns(df, metrics.keys())
        self._dataframes[df_name] = pd.concat([df, pd.DataFrame([metrics])], ignore_index=True)

    def _process_batch(self, param, df_name):
        now = time.time()
        if param.eval_metric is not None:
            metrics = dict(param.eval_metric.get_name_value())
            param.eval_metric.reset()
        else:
            metrics = {}
        speed = self.frequent * self.batch_size / max(now - self.last_time, 1e-9)
        metrics.update({"batches_per_sec": speed / self.batch_size, "records_per_sec": speed,
                        "elapsed": self.elapsed(), "minibatch_count": param.nbatch, "epoch": param.epoch})
        self.append_metrics(metrics, df_name)
        self.last_time = now

    def train_cb(self, param):
        if param.nbatch % self.frequent == 0:
            self._process_batch(param, "train")

    def eval_cb(self, param):
        self._process_batch(param, "eval")

    def epoch_cb(self, *_a):
        now = datetime.datetime.now()
        self.append_metrics({"elapsed": self.elapsed(), "epoch_time": now - self.last_epoch_time}, "epoch")
        self.last_epoch_time = now

    def callback_args(self):
        return {"batch_end_callback": self.train_cb, "eval_end_callback": self.eval_cb, "epoch_end_callback": self.epoch_cb}


class LiveLearningCurve(PandasLogger):
    """Collects the named metric for a live learning curve (plotting is left to the notebook front-end)."""

    def __init__(self, metric_name, display_freq=10, frequent=50, batch_size=1):
        super(LiveLearningCurve, self).__init__(batch_size, frequent)
        self.metric_name, self.display_freq = metric_name, display_freq
