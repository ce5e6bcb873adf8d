"""Legacy data-parallel executor manager used by ``FeedForward`` (reference: python/mxnet/executor_manager.py).
Thin layer over ``module.executor_group.DataParallelExecutorGroup``: batch slicing by work load, per-device executors,
forward / backward / metric update and parameter copies."""
import logging

import numpy as np

from .module.executor_group import DataParallelExecutorGroup


def _split_input_slice(batch_size, work_load_list):
    """Slices of a batch proportional to the work loads (executor_manager.py:31-66)."""
    total = sum(work_load_list)
    nums = [round(w * batch_size / total) for w in work_load_list]
    if sum(nums) < batch_size:
        nums[-1] += batch_size - sum(nums)
    slices, end = [], 0
    for n in nums:
        begin = int(min(end, batch_size))
        end = int(min(begin + n, batch_size))
        if begin >= end:
            raise ValueError("Too many slices. Some splits are empty.")
        slices.append(slice(begin, end))
    return slices


def _check_arguments(symbol):
    arg_names = symbol.list_arguments()
    if len(set(arg_names)) != len(arg_names):
        raise ValueError("Find duplicated argument name; please make the weight name non-duplicated (using name arguments), "
                         "arguments are %s" % str(arg_names))
    aux_names = symbol.list_auxiliary_states()
    if len(set(aux_names)) != len(aux_names):
        raise ValueError("Find duplicated auxiliary param name, auxiliary params are %s" % str(aux_names))


class DataParallelExecutorManager(object):
    def __init__(self, symbol, ctx, train_data, arg_names, param_names, aux_names, work_load_list=None, logger=None,
                 sym_gen=None):
        self.logger = logger or logging
        num_device = len(ctx)
        self.logger.info("Start training with %s", str(ctx))
        work_load_list = work_load_list or [1] * num_device
        assert isinstance(work_load_list, list) and len(work_load_list) == num_device, "Invalid settings for work load."
        self.slices = _split_input_slice(train_data.batch_size, work_load_list)
        self.arg_names, self.param_names, self.aux_names, self.ctx = arg_names, param_names, aux_names, ctx
        _check_arguments(symbol)
        self.symbol, self.sym_gen = symbol, sym_gen
        self.execgrp = DataParallelExecutorGroup(symbol, ctx, work_load_list, train_data.provide_data, train_data.provide_label,
                                                 param_names, for_training=True, inputs_need_grad=False)
        self.curr_execgrp = self.execgrp

    def install_monitor(self, monitor):
        if self.sym_gen is not None:
            raise NotImplementedError("Monitoring is not implemented for bucketing")
        self.execgrp.install_monitor(monitor)

    def set_params(self, arg_params, aux_params):
        self.execgrp.set_params(arg_params, aux_params)

    def copy_to(self, arg_params, aux_params):
        self.execgrp.get_params(arg_params, aux_params)

    @property
    def param_arrays(self):
        return self.execgrp.param_arrays

    @property
    def grad_arrays(self):
        return self.execgrp.grad_arrays

    @property
    def aux_arrays(self):
        return self.execgrp.aux_arrays

    def load_data_batch(self, data_batch):
        self._batch = data_batch

    def forward(self, is_train=False):
        self.curr_execgrp.forward(self._batch, is_train=is_train)

    def backward(self):
        self.curr_execgrp.backward()

    def update_metric(self, metric, labels, pre_sliced=False):
        self.curr_execgrp.update_metric(metric, labels, pre_sliced)
