"""Evaluation metrics (reference: python/mxnet/metric.py:363-1392).

Provenance: adapted from Apache MXNet (incubating) 1.3.1 ``python/mxnet/metric.py`` (Apache License 2.0).  The public API, argument
names and control flow follow the original so that user scripts keep working; the tensor maths runs on torch /
the mxb200 kernels.  This file is API-compatibility glue, not original B200 work ("Provenance").
"""
import math
from collections import OrderedDict

import numpy
import torch

from .ndarray import NDArray

__all__ = ["EvalMetric", "CompositeEvalMetric", "Accuracy", "TopKAccuracy", "F1", "MCC", "Perplexity", "MAE", "MSE",
           "RMSE", "CrossEntropy", "NegativeLogLikelihood", "PearsonCorrelation", "Loss", "Torch", "Caffe", "CustomMetric",
           "np", "create", "register", "check_label_shapes"]

_REG = {}


def register(klass, *names):
    for n in (klass.__name__.lower(),) + names:
        _REG[n] = klass
    return klass


def alias(*names):
    def deco(klass):
        return register(klass, *names)
    return deco


def create(metric, *args, **kwargs):
    if callable(metric) and not isinstance(metric, type):
        return CustomMetric(metric, *args, **kwargs)
    if isinstance(metric, EvalMetric):
        return metric
    if isinstance(metric, list):
        composite = CompositeEvalMetric()
        for child in metric:
            composite.add(create(child, *args, **kwargs))
        return composite
    if isinstance(metric, dict):
        cfg = dict(metric)
        return create(cfg.pop("metric"), *args, **dict(cfg, **kwargs))
    if isinstance(metric, str):
        # serialized forms produced by EvalMetric.get_config() + json (registry.py:get_create_func of the reference)
        if metric.startswith("{"):
            import json
            return create(json.loads(metric), *args, **kwargs)
        if metric.startswith("["):
            import json
            name, cfg = json.loads(metric)
            return create(name, *args, **dict(cfg, **kwargs))
        if metric.lower() not in _REG:
            raise ValueError("Metric must be either callable or in %s; got %r" % (sorted(_REG), metric))
        return _REG[metric.lower()](*args, **kwargs)
    raise TypeError("metric should be a str, callable, EvalMetric or list")


def check_label_shapes(labels, preds, wrap=False, shape=False):
    if not shape:
        label_shape, pred_shape = len(labels), len(preds)
    else:
        label_shape, pred_shape = labels.shape, preds.shape
    if label_shape != pred_shape:
        raise ValueError("Shape of labels {} does not match shape of predictions {}".format(label_shape, pred_shape))
    if wrap:
        if isinstance(labels, NDArray):
            labels = [labels]
        if isinstance(preds, NDArray):
            preds = [preds]
    return labels, preds


def _np(x):
    return x.asnumpy() if isinstance(x, NDArray) else numpy.asarray(x)


class EvalMetric(object):
    def __init__(self, name, output_names=None, label_names=None, **kwargs):
        self.name = str(name)
        self.output_names = output_names
        self.label_names = label_names
        self._kwargs = kwargs
        self.reset()

    def __str__(self):
        return "EvalMetric: {}".format(dict(self.get_name_value()))

    def get_config(self):
        config = self._kwargs.copy()
        config.update({"metric": self.__class__.__name__, "name": self.name, "output_names": self.output_names,
                       "label_names": self.label_names})
        return config

    def update_dict(self, label, pred):
        if self.output_names is not None:
            pred = [pred[name] for name in self.output_names]
        else:
            pred = list(pred.values())
        if self.label_names is not None:
            label = [label[name] for name in self.label_names]
        else:
            label = list(label.values())
        self.update(label, pred)

    def update(self, labels, preds):
        raise NotImplementedError()

    def reset(self):
        self.num_inst = 0
        self.sum_metric = 0.0

    def get(self):
        if self.num_inst == 0:
            return (self.name, float("nan"))
        return (self.name, self.sum_metric / self.num_inst)

    def get_name_value(self):
        name, value = self.get()
        if not isinstance(name, list):
            name = [name]
        if not isinstance(value, list):
            value = [value]
        return list(zip(name, value))


@register
class CompositeEvalMetric(EvalMetric):
    def __init__(self, metrics=None, name="composite", output_names=None, label_names=None):
        super(CompositeEvalMetric, self).__init__(name, output_names=output_names, label_names=label_names)
        if metrics is None:
            metrics = []
        self.metrics = [create(i) for i in metrics]

    def add(self, metric):
        self.metrics.append(create(metric))

    def get_metric(self, index):
        try:
            return self.metrics[index]
        except IndexError:
            return ValueError("Metric index {} is out of range 0 and {}".format(index, len(self.metrics)))

    def update_dict(self, labels, preds):
        if self.label_names is not None:
            labels = OrderedDict([i for i in labels.items() if i[0] in self.label_names])
        if self.output_names is not None:
            preds = OrderedDict([i for i in preds.items() if i[0] in self.output_names])
        for metric in self.metrics:
            metric.update_dict(labels, preds)

    def update(self, labels, preds):
        for metric in self.metrics:
            metric.update(labels, preds)

    def reset(self):
        try:
            for metric in self.metrics:
                metric.reset()
        except AttributeError:
            pass

    def get(self):
        names, values = [], []
        for metric in self.metrics:
            name, value = metric.get()
            if isinstance(name, str):
                name = [name]
            if isinstance(value, (float, int, numpy.generic)):
                value = [value]
            names.extend(name)
            values.extend(value)
        return (names, values)


@alias("acc")
class Accuracy(EvalMetric):
    def __init__(self, axis=1, name="accuracy", output_names=None, label_names=None):
        super(Accuracy, self).__init__(name, axis=axis, output_names=output_names, label_names=label_names)
        self.axis = axis

    def update(self, labels, preds):
        labels, preds = check_label_shapes(labels, preds, True)
        for label, pred in zip(labels, preds):
            p, l = pred._t.detach(), label._t.detach()
            if p.shape != l.shape:
                p = torch.argmax(p.float(), dim=self.axis)
            p = p.reshape(-1).to(torch.int64)
            l = l.reshape(-1).to(torch.int64).to(p.device)
            check_label_shapes(l, p)
            self.sum_metric += float((p == l).sum().item())
            self.num_inst += int(p.numel())


@alias("top_k_accuracy", "top_k_acc")
class TopKAccuracy(EvalMetric):
    def __init__(self, top_k=1, name="top_k_accuracy", output_names=None, label_names=None):
        super(TopKAccuracy, self).__init__(name, top_k=top_k, output_names=output_names, label_names=label_names)
        self.top_k = top_k
        assert self.top_k > 1, "Please use Accuracy if top_k is no more than 1"
        self.name += "_%d" % self.top_k

    def update(self, labels, preds):
        labels, preds = check_label_shapes(labels, preds, True)
        for label, pred in zip(labels, preds):
            p, l = pred._t.detach().float(), label._t.detach().reshape(-1).to(torch.int64)
            assert p.dim() <= 2, "Predictions should be no more than 2 dims"
            if p.dim() == 1:
                self.sum_metric += float((p.to(torch.int64) == l.to(p.device)).sum().item())
            else:
                k = min(self.top_k, p.shape[1])
                top = torch.topk(p, k, dim=1).indices
                self.sum_metric += float((top == l.to(p.device)[:, None]).any(dim=1).sum().item())
            self.num_inst += int(l.numel())


class _BinaryClassificationMetrics(object):
    def __init__(self):
        self.reset_stats()

    def update_binary_stats(self, label, pred):
        pred, label = _np(pred), _np(label).astype("int32")
        pred_label = numpy.argmax(pred, axis=1)
        check_label_shapes(label, pred)
        if len(numpy.unique(label)) > 2:
            raise ValueError("%s currently only supports binary classification." % self.__class__.__name__)
        self.true_positives += ((pred_label == 1) & (label == 1)).sum()
        self.false_positives += ((pred_label == 1) & (label == 0)).sum()
        self.false_negatives += ((pred_label == 0) & (label == 1)).sum()
        self.true_negatives += ((pred_label == 0) & (label == 0)).sum()

    @property
    def precision(self):
        d = self.true_positives + self.false_positives
        return self.true_positives / float(d) if d > 0 else 0.0

    @property
    def recall(self):
        d = self.true_positives + self.false_negatives
        return self.true_positives / float(d) if d > 0 else 0.0

    @property
    def fscore(self):
        if self.precision + self.recall > 0:
            return 2 * self.precision * self.recall / (self.precision + self.recall)
        return 0.0

    @property
    def matthewscc(self):
        tp, fp, fn, tn = self.true_positives, self.false_positives, self.false_negatives, self.true_negatives
        terms = [(tp + fp), (tp + fn), (tn + fp), (tn + fn)]
        denom = 1.0
        for t in filter(lambda t: t != 0.0, terms):
            denom *= t
        return ((tp * tn) - (fp * fn)) / math.sqrt(denom)

    @property
    def total_examples(self):
        return self.false_negatives + self.false_positives + self.true_negatives + self.true_positives

    def reset_stats(self):
        self.false_positives = self.false_negatives = self.true_positives = self.true_negatives = 0


@register
class F1(EvalMetric):
    def __init__(self, name="f1", output_names=None, label_names=None, average="macro"):
        self.average = average
        self.metrics = _BinaryClassificationMetrics()
        EvalMetric.__init__(self, name=name, output_names=output_names, label_names=label_names)

    def update(self, labels, preds):
        labels, preds = check_label_shapes(labels, preds, True)
        for label, pred in zip(labels, preds):
            self.metrics.update_binary_stats(label, pred)
        if self.average == "macro":
            self.sum_metric += self.metrics.fscore
            self.num_inst += 1
            self.metrics.reset_stats()
        else:
            self.sum_metric = self.metrics.fscore * self.metrics.total_examples
            self.num_inst = self.metrics.total_examples

    def reset(self):
        self.sum_metric = 0.
        self.num_inst = 0.
        self.metrics.reset_stats()


@register
class MCC(EvalMetric):
    def __init__(self, name="mcc", output_names=None, label_names=None, average="macro"):
        self._average = average
        self._metrics = _BinaryClassificationMetrics()
        EvalMetric.__init__(self, name=name, output_names=output_names, label_names=label_names)

    def update(self, labels, preds):
        labels, preds = check_label_shapes(labels, preds, True)
        for label, pred in zip(labels, preds):
            self._metrics.update_binary_stats(label, pred)
        if self._average == "macro":
            self.sum_metric += self._metrics.matthewscc
            self.num_inst += 1
            self._metrics.reset_stats()
        else:
            self.sum_metric = self._metrics.matthewscc * self._metrics.total_examples
            self.num_inst = self._metrics.total_examples

    def reset(self):
        self.sum_metric = 0.
        self.num_inst = 0.
        self._metrics.reset_stats()


@register
class Perplexity(EvalMetric):
    def __init__(self, ignore_label, axis=-1, name="perplexity", output_names=None, label_names=None):
        super(Perplexity, self).__init__(name, ignore_label=ignore_label, output_names=output_names, label_names=label_names)
        self.ignore_label = ignore_label
        self.axis = axis

    def update(self, labels, preds):
        assert len(labels) == len(preds)
        loss, num = 0., 0
        for label, pred in zip(labels, preds):
            l = label._t.detach().reshape(-1).to(torch.int64)
            p = pred._t.detach().float()
            p = p.reshape(-1, p.shape[-1]) if self.axis in (-1, p.dim() - 1) else p.movedim(self.axis, -1).reshape(-1, p.shape[self.axis])
            probs = p.gather(1, l.to(p.device)[:, None]).squeeze(1)
            if self.ignore_label is not None:
                ignore = (l == self.ignore_label).to(probs.device)
                num -= int(ignore.sum().item())
                probs = probs * (~ignore).float() + ignore.float()
            loss -= float(torch.log(torch.clamp(probs, min=1e-10)).sum().item())
            num += int(probs.numel())
        self.sum_metric += loss
        self.num_inst += num

    def get(self):
        if self.num_inst == 0:
            return (self.name, float("nan"))
        return (self.name, math.exp(self.sum_metric / self.num_inst))


def _regression(name_, fn):
    class _M(EvalMetric):
        def __init__(self, name=name_, output_names=None, label_names=None):
            super(_M, self).__init__(name, output_names=output_names, label_names=label_names)

        def update(self, labels, preds):
            labels, preds = check_label_shapes(labels, preds, True)
            for label, pred in zip(labels, preds):
                l, p = _np(label), _np(pred)
                if len(l.shape) == 1:
                    l = l.reshape(l.shape[0], 1)
                if len(p.shape) == 1:
                    p = p.reshape(p.shape[0], 1)
                self.sum_metric += fn(l, p)
                self.num_inst += 1
    _M.__name__ = name_.upper()
    return _M


MAE = register(_regression("mae", lambda l, p: numpy.abs(l - p).mean()))
MSE = register(_regression("mse", lambda l, p: ((l - p) ** 2.0).mean()))
RMSE = register(_regression("rmse", lambda l, p: numpy.sqrt(((l - p) ** 2.0).mean())))


@alias("ce")
class CrossEntropy(EvalMetric):
    def __init__(self, eps=1e-12, name="cross-entropy", output_names=None, label_names=None):
        super(CrossEntropy, self).__init__(name, eps=eps, output_names=output_names, label_names=label_names)
        self.eps = eps

    def update(self, labels, preds):
        labels, preds = check_label_shapes(labels, preds, True)
        for label, pred in zip(labels, preds):
            l, p = _np(label).ravel(), _np(pred)
            assert l.shape[0] == p.shape[0]
            prob = p[numpy.arange(l.shape[0]), numpy.int64(l)]
            self.sum_metric += (-numpy.log(prob + self.eps)).sum()
            self.num_inst += l.shape[0]


@alias("nll_loss")
class NegativeLogLikelihood(CrossEntropy):
    def __init__(self, eps=1e-12, name="nll-loss", output_names=None, label_names=None):
        super(NegativeLogLikelihood, self).__init__(eps, name, output_names, label_names)


@alias("pearsonr")
class PearsonCorrelation(EvalMetric):
    def __init__(self, name="pearsonr", output_names=None, label_names=None):
        super(PearsonCorrelation, self).__init__(name, output_names=output_names, label_names=label_names)

    def update(self, labels, preds):
        labels, preds = check_label_shapes(labels, preds, True)
        for label, pred in zip(labels, preds):
            check_label_shapes(label, pred, False, True)
            self.sum_metric += numpy.corrcoef(_np(pred).ravel(), _np(label).ravel())[0, 1]
            self.num_inst += 1


@register
class Loss(EvalMetric):
    """Mean of the values passed as predictions (used to track the training loss)."""

    def __init__(self, name="loss", output_names=None, label_names=None):
        super(Loss, self).__init__(name, output_names=output_names, label_names=label_names)

    def update(self, _, preds):
        if isinstance(preds, NDArray):
            preds = [preds]
        for pred in preds:
            self.sum_metric += float(pred._t.detach().float().sum().item())
            self.num_inst += pred.size


@register
class Torch(Loss):
    def __init__(self, name="torch", output_names=None, label_names=None):
        super(Torch, self).__init__(name, output_names=output_names, label_names=label_names)


@register
class Caffe(Loss):
    def __init__(self, name="caffe", output_names=None, label_names=None):
        super(Caffe, self).__init__(name, output_names=output_names, label_names=label_names)


@register
class CustomMetric(EvalMetric):
    def __init__(self, feval, name=None, allow_extra_outputs=False, output_names=None, label_names=None):
        if name is None:
            name = feval.__name__
            if name.find("<") != -1:
                name = "custom(%s)" % name
        super(CustomMetric, self).__init__(name, feval=feval, allow_extra_outputs=allow_extra_outputs,
                                           output_names=output_names, label_names=label_names)
        self._feval = feval
        self._allow_extra_outputs = allow_extra_outputs

    def update(self, labels, preds):
        if not self._allow_extra_outputs:
            labels, preds = check_label_shapes(labels, preds, True)
        for pred, label in zip(preds, labels):
            reval = self._feval(_np(label), _np(pred))
            if isinstance(reval, tuple):
                (sum_metric, num_inst) = reval
                self.sum_metric += sum_metric
                self.num_inst += num_inst
            else:
                self.sum_metric += reval
                self.num_inst += 1

    def get_config(self):
        raise NotImplementedError("CustomMetric cannot be serialized")


def np(numpy_feval, name=None, allow_extra_outputs=False):  # noqa: F811  (reference name: mx.metric.np)
    def feval(label, pred):
        return numpy_feval(label, pred)
    feval.__name__ = numpy_feval.__name__
    return CustomMetric(feval, name, allow_extra_outputs)
