"""DataParallelExecutorGroup: one executor per context, batches sliced along the batch axis
(reference: python/mxnet/module/executor_group.py:143-700)."""
import logging

import numpy as np
import torch

from ..base import MXNetError
from ..io import DataDesc
from .. import ndarray as nd
from ..ndarray import NDArray
from ..executor import Executor

__all__ = ["DataParallelExecutorGroup"]


def _split_input_slice(batch_size, work_load_list):
    """Slices proportional to the workloads (executor_manager._split_input_slice)."""
    total = sum(work_load_list)
    batch_num_list = [round(w * batch_size / total) for w in work_load_list]
    batch_num_sum = sum(batch_num_list)
    if batch_num_sum < batch_size:
        batch_num_list[-1] += batch_size - batch_num_sum
    slices, end = [], 0
    for n in batch_num_list:
        begin = int(min(end, batch_size))
        end = int(min(begin + n, batch_size))
        if begin >= end:
            raise ValueError("Too many slices. Some splits are empty.")
        slices.append(slice(begin, end))
    return slices


class DataParallelExecutorGroup(object):
    def __init__(self, symbol, contexts, workload, data_shapes, label_shapes, param_names, for_training,
                 inputs_need_grad, shared_group=None, logger=logging, fixed_param_names=None, grad_req="write",
                 state_names=None, group2ctxs=None):
        self.param_names = param_names
        self.arg_names = symbol.list_arguments()
        self.aux_names = symbol.list_auxiliary_states()
        self.symbol = symbol
        self.contexts = contexts
        self.workload = workload
        self.group2ctxs = group2ctxs
        self.for_training = for_training
        self.inputs_need_grad = inputs_need_grad
        self.logger = logger
        self.fixed_param_names = fixed_param_names or []
        self.state_names = state_names or []
        data_names = [x[0] if not isinstance(x, DataDesc) else x.name for x in data_shapes]
        label_names = [x[0] if not isinstance(x, DataDesc) else x.name for x in (label_shapes or [])]
        if isinstance(grad_req, str):
            self.grad_req = {}
            for k in self.arg_names:
                if k in self.param_names:
                    self.grad_req[k] = "null" if k in self.fixed_param_names or not for_training else grad_req
                elif k in data_names:
                    self.grad_req[k] = grad_req if self.inputs_need_grad else "null"
                else:
                    self.grad_req[k] = "null"
        else:
            self.grad_req = {k: "null" for k in self.arg_names}
            self.grad_req.update(grad_req if isinstance(grad_req, dict) else dict(zip(self.arg_names, grad_req)))
        self.execs = []
        self.data_arrays = self.label_arrays = self.param_arrays = self.grad_arrays = self.aux_arrays = None
        self.batch_size = None
        self.slices = None
        self.data_shapes = self.label_shapes = None
        self.data_names = self.label_names = None
        self.output_names = self.symbol.list_outputs()
        self.num_outputs = len(self.output_names)
        self.bind_exec(data_shapes, label_shapes, shared_group)

    def decide_slices(self, data_shapes):
        assert len(data_shapes) > 0
        major_axis = [DataDesc.get_batch_axis(getattr(x, "layout", "NCHW")) for x in data_shapes]
        for (name, shape), axis in zip([(x[0], x[1]) for x in data_shapes], major_axis):
            if axis == -1:
                continue
            batch_size = shape[axis]
            if self.batch_size is not None:
                assert batch_size == self.batch_size, ("all data must have the same batch size: " +
                                                       ("batch_size = %d, but " % self.batch_size) +
                                                       ("%s has shape %s" % (name, shape)))
            else:
                self.batch_size = batch_size
                self.slices = _split_input_slice(self.batch_size, self.workload)
        return major_axis

    def bind_exec(self, data_shapes, label_shapes, shared_group=None, reshape=False):
        self.batch_size = None
        self.data_layouts = self.decide_slices(data_shapes)
        if label_shapes is not None:
            self.label_layouts = self.decide_slices(label_shapes)
        self.data_shapes = data_shapes
        self.label_shapes = label_shapes
        self.data_names = [i[0] for i in data_shapes]
        self.label_names = [i[0] for i in label_shapes] if label_shapes is not None else []
        old_execs = self.execs
        self.execs = []
        # input arrays are pooled per device and shared between the executor groups of a bucketing module: a bucket whose
        # inputs fit into an array that exists already gets a view of it (executor_group.py shared_data_arrays)
        if shared_group is not None and getattr(shared_group, "shared_data_arrays", None) is not None:
            self.shared_data_arrays = shared_group.shared_data_arrays
        elif getattr(self, "shared_data_arrays", None) is None or reshape:
            self.shared_data_arrays = [{} for _ in self.contexts]
        for i, ctx in enumerate(self.contexts):
            shapes = {}
            for desc, axis in zip(data_shapes, self.data_layouts):
                shapes[desc[0]] = self._sliced_shape(desc[1], i, axis)
            if label_shapes is not None:
                for desc, axis in zip(label_shapes, self.label_layouts):
                    shapes[desc[0]] = self._sliced_shape(desc[1], i, axis)
            type_dict = {}
            for desc in list(data_shapes) + list(label_shapes or []):
                if isinstance(desc, DataDesc):
                    type_dict[desc.name] = desc.dtype
            ex = Executor.simple_bind(self.symbol, ctx, self.grad_req, type_dict, self._group2ctx_of(i), shapes)
            if reshape and old_execs:
                ex.copy_params_from({n: old_execs[i].arg_dict[n] for n in self.param_names},
                                    old_execs[i].aux_dict, allow_extra_params=True)
            elif shared_group is not None:
                # share parameter arrays with the shared module (bucketing)
                for n in self.param_names:
                    if n in shared_group.execs[i].arg_dict:
                        ex.arg_dict[n] = shared_group.execs[i].arg_dict[n]
                        if n in shared_group.execs[i].grad_dict:
                            ex.grad_dict[n] = shared_group.execs[i].grad_dict[n]
                for n in self.aux_names:
                    if n in shared_group.execs[i].aux_dict:
                        ex.aux_dict[n] = shared_group.execs[i].aux_dict[n]
                ex.arg_arrays = [ex.arg_dict[n] for n in self.arg_names]
                ex.grad_arrays = [ex.grad_dict.get(n) for n in self.arg_names]
                ex.aux_arrays = [ex.aux_dict[n] for n in self.aux_names]
            pool = self.shared_data_arrays[i]
            if shared_group is not None and not (reshape and old_execs):
                # arrays this group had to create itself (parameters the shared group does not have) join the pool too
                for n in self.arg_names:
                    if n not in shared_group.execs[i].arg_dict and n not in self.data_names + self.label_names:
                        pool.setdefault(n, ex.arg_dict[n])
            for n in self.data_names + self.label_names:
                if n not in ex.arg_dict or ex._grad_req.get(n, "null") != "null":
                    continue
                arr = ex.arg_dict[n]
                have = pool.get(n)
                if have is not None and have.size >= arr.size and have.dtype == arr.dtype and have.context == arr.context \
                        and have._t.is_contiguous():
                    view = have if have.shape == arr.shape else NDArray(have._t.reshape(-1)[:arr.size].view(arr.shape), have.context)
                    ex.arg_dict[n] = view
                else:
                    pool[n] = arr
            ex.arg_arrays = [ex.arg_dict[n] for n in self.arg_names]
            ex._rebind_arrays()
            self.execs.append(ex)
        self._collect_arrays()

    def _group2ctx_of(self, i):
        """ctx_group -> context map of the i-th executor (executor_group.py:_prepare_group2ctxs of the reference): one dict
        for every device, a list of dicts (one per device), or a dict whose values are per-device lists."""
        g = self.group2ctxs
        if not g:
            return None
        if isinstance(g, (list, tuple)):
            assert len(g) == len(self.contexts), "length of group2ctxs should be %d" % len(self.contexts)
            return g[i]
        out = {}
        for k, v in g.items():
            if isinstance(v, (list, tuple)):
                assert len(v) in (1, len(self.contexts)), "length of group2ctxs[%s] should be %d or 1" % (k, len(self.contexts))
                out[k] = v[i] if len(v) > 1 else v[0]
            else:
                out[k] = v
        return out

    def _sliced_shape(self, shape, i, axis):
        shape = list(shape)
        if axis >= 0:
            shape[axis] = self.slices[i].stop - self.slices[i].start
        return tuple(shape)

    def reshape(self, data_shapes, label_shapes):
        if data_shapes == self.data_shapes and label_shapes == self.label_shapes:
            return
        self.bind_exec(data_shapes, label_shapes, reshape=True)

    def _collect_arrays(self):
        self.data_arrays = [[(self.slices[i], e.arg_dict[name]) for i, e in enumerate(self.execs)] for name in self.data_names]
        self.state_arrays = [[e.arg_dict[name] for e in self.execs] for name in self.state_names]
        if self.label_shapes is not None:
            self.label_arrays = [[(self.slices[i], e.arg_dict[name]) for i, e in enumerate(self.execs)]
                                 for name in self.label_names if name in self.execs[0].arg_dict]
        else:
            self.label_arrays = None
        self.param_arrays = [[e.arg_dict[name] for e in self.execs] for name in self.param_names]
        if self.for_training:
            self.grad_arrays = [[e.grad_dict.get(name) for e in self.execs] for name in self.param_names]
        else:
            self.grad_arrays = None
        if self.inputs_need_grad:
            self.input_grad_arrays = [[e.grad_dict.get(name) for e in self.execs] for name in self.data_names]
        else:
            self.input_grad_arrays = None
        self.aux_arrays = [[e.aux_dict[name] for e in self.execs] for name in self.aux_names]

    def set_params(self, arg_params, aux_params, allow_extra=False):
        for exec_ in self.execs:
            exec_.copy_params_from(arg_params, aux_params, allow_extra_params=allow_extra)

    def get_params(self, arg_params, aux_params):
        """Average the per-device copies into arg_params / aux_params on the host
        (executor_group.py:415-434) -- also how BN moving statistics are merged within a box."""
        for name, block in zip(self.param_names, self.param_arrays):
            weight = sum(w._t.detach().to("cpu").float() for w in block) / len(block)
            arg_params[name] = NDArray(weight.to(block[0]._t.dtype)) if name not in arg_params else arg_params[name]
            with torch.no_grad():
                arg_params[name]._t.copy_(weight.to(arg_params[name]._t.dtype))
        for name, block in zip(self.aux_names, self.aux_arrays):
            weight = sum(w._t.detach().to("cpu").float() for w in block) / len(block)
            aux_params[name] = NDArray(weight.to(block[0]._t.dtype)) if name not in aux_params else aux_params[name]
            with torch.no_grad():
                aux_params[name]._t.copy_(weight.to(aux_params[name]._t.dtype))

    def _load(self, batch_arrays, targets, layouts):
        for d_src, d_targets, axis in zip(batch_arrays, targets, layouts):
            for slice_idx, d_dst in d_targets:
                src = d_src._t if isinstance(d_src, NDArray) else torch.as_tensor(np.asarray(d_src))
                if axis >= 0 and len(d_targets) > 1:
                    src = src.narrow(axis, slice_idx.start, slice_idx.stop - slice_idx.start)
                elif axis >= 0 and src.shape[axis] != d_dst.shape[axis]:
                    src = src.narrow(axis, slice_idx.start, slice_idx.stop - slice_idx.start)
                with torch.no_grad():
                    d_dst._t.copy_(src.to(d_dst._t.dtype), non_blocking=True)

    def forward(self, data_batch, is_train=None):
        if is_train is None:
            is_train = self.for_training
        self._load(data_batch.data, self.data_arrays, self.data_layouts)
        if self.label_arrays is not None and data_batch.label:
            self._load(data_batch.label, self.label_arrays, self.label_layouts)
        for exec_ in self.execs:
            exec_.forward(is_train=is_train)

    def get_output_shapes(self):
        outputs = self.execs[0].outputs
        shapes = [out.shape for out in outputs]
        concat_shapes = []
        for key, the_shape in zip(self.symbol.list_outputs(), shapes):
            the_shape = list(the_shape)
            if the_shape:
                the_shape[0] = self.batch_size
            concat_shapes.append((key, tuple(the_shape)))
        return concat_shapes

    def get_outputs(self, merge_multi_context=True, begin=0, end=None):
        if end is None:
            end = self.num_outputs
        outputs = [[exec_.outputs[i] for exec_ in self.execs] for i in range(begin, end)]
        if merge_multi_context:
            axes = self._output_batch_axes()[begin:end]
            return [_merge(x, axis=ax) for x, ax in zip(outputs, axes)]
        return outputs

    def _output_batch_axes(self):
        """Batch axis of every output from its ``__layout__`` attribute (executor_group.py:241-244 of the reference); 0 when
        the symbol does not say."""
        axes = getattr(self, "_out_axes", None)
        if axes is None:
            axes = []
            for i in range(self.num_outputs):
                lay = self.symbol[i].attr("__layout__")
                ax = DataDesc.get_batch_axis(lay) if lay else 0
                axes.append(0 if ax is None or ax < 0 else ax)
            self._out_axes = axes
        return axes

    def get_states(self, merge_multi_context=True):
        return [_merge(x) for x in self.state_arrays] if merge_multi_context else self.state_arrays

    def set_states(self, states=None, value=None):
        if states is not None:
            for i, (d_dst, src) in enumerate(zip(self.state_arrays, states)):
                for dst, s in zip(d_dst, src if isinstance(src, (list, tuple)) else [src] * len(d_dst)):
                    dst[:] = s
        else:
            for d_dst in self.state_arrays:
                for dst in d_dst:
                    dst[:] = value

    def get_input_grads(self, merge_multi_context=True):
        assert self.inputs_need_grad
        if merge_multi_context:
            return [_merge(x) for x in self.input_grad_arrays]
        return self.input_grad_arrays

    def backward(self, out_grads=None):
        assert self.for_training, "re-bind with for_training=True to run backward"
        if out_grads is None:
            out_grads = []
        for i, (exec_, islice) in enumerate(zip(self.execs, self.slices)):
            out_grads_slice = []
            for grad, ax in zip(out_grads, self._output_batch_axes()):
                g = grad._t
                if len(self.execs) > 1:
                    g = g.narrow(ax, islice.start, islice.stop - islice.start)
                out_grads_slice.append(NDArray(g.to(exec_._ctx.torch_device), exec_._ctx))
            exec_.backward(out_grads=out_grads_slice)
        if self.for_training:
            self.grad_arrays = [[e.grad_dict.get(name) for e in self.execs] for name in self.param_names]

    def update_metric(self, eval_metric, labels, pre_sliced=False):
        for current_exec, (texec, islice) in enumerate(zip(self.execs, self.slices)):
            if not pre_sliced:
                labels_slice = []
                for label in labels:
                    if len(self.execs) > 1:
                        labels_slice.append(NDArray(label._t.narrow(0, islice.start, islice.stop - islice.start), label.context))
                    else:
                        labels_slice.append(label)
            else:
                labels_slice = labels[current_exec]
            labels_ = dict(zip(self.label_names, labels_slice))
            preds = dict(zip(self.output_names, texec.outputs))
            eval_metric.update_dict(labels_, preds)

    def install_monitor(self, mon):
        for exe in self.execs:
            mon.install(exe)


def _merge(tensors, axis=0):
    if len(tensors) == 1:
        return tensors[0]
    dev = tensors[0]._t.device
    return NDArray(torch.cat([t._t.to(dev) for t in tensors], dim=axis), tensors[0].context)
