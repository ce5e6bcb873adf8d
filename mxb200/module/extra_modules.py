"""BucketingModule, SequentialModule, PythonModule (reference: python/mxnet/module/{bucketing,sequential,python}_module.py)."""
import logging

from .. import ndarray as nd
from ..initializer import Uniform
from .base_module import BaseModule
from .module import Module


class BucketingModule(BaseModule):
    """One Module per bucket key sharing parameters with the default bucket."""

    def __init__(self, sym_gen, default_bucket_key=None, logger=logging, context=None, work_load_list=None,
                 fixed_param_names=None, state_names=None, group2ctxs=None, compression_params=None):
        super(BucketingModule, self).__init__(logger=logger)
        assert default_bucket_key is not None
        self._default_bucket_key = default_bucket_key
        self._sym_gen = sym_gen
        self._context, self._work_load_list = context, work_load_list
        self._fixed_param_names, self._state_names = fixed_param_names, state_names
        self._group2ctxs, self._compression_params = group2ctxs, compression_params
        self._buckets = {}
        self._curr_module = None
        self._curr_bucket_key = None
        self._params_dirty = False
        self._grad_req = "write"

    def _gen(self, key):
        sym, dnames, lnames = self._sym_gen(key)
        m = Module(sym, dnames, lnames, logger=self.logger, context=self._context,
                   work_load_list=self._work_load_list, fixed_param_names=self._fixed_param_names,
                   state_names=self._state_names, group2ctxs=self._group2ctxs,
                   compression_params=self._compression_params)
        m._no_fuse = True     # buckets share parameter arrays between executors; keep the per-key update path
        return m

    @property
    def data_names(self):
        return self._curr_module.data_names if self.binded else self._sym_gen(self._default_bucket_key)[1]

    @property
    def output_names(self):
        return self._curr_module.output_names if self.binded else self._sym_gen(self._default_bucket_key)[0].list_outputs()

    @property
    def data_shapes(self):
        return self._curr_module.data_shapes

    @property
    def label_shapes(self):
        return self._curr_module.label_shapes

    @property
    def output_shapes(self):
        return self._curr_module.output_shapes

    @property
    def symbol(self):
        return self._curr_module.symbol

    def get_params(self):
        self._curr_module._params_dirty = self._params_dirty
        params = self._curr_module.get_params()
        self._params_dirty = False
        return params

    def set_params(self, arg_params, aux_params, allow_missing=False, force_init=True, allow_extra=False):
        self._curr_module.set_params(arg_params, aux_params, allow_missing, force_init, allow_extra)
        self._params_dirty = True
        self.params_initialized = True

    def init_params(self, initializer=Uniform(0.01), arg_params=None, aux_params=None, allow_missing=False,
                    force_init=False, allow_extra=False):
        if self.params_initialized and not force_init:
            return
        self._curr_module.init_params(initializer, arg_params, aux_params, allow_missing, force_init, allow_extra)
        self._params_dirty = False
        self.params_initialized = True

    def bind(self, data_shapes, label_shapes=None, for_training=True, inputs_need_grad=False, force_rebind=False,
             shared_module=None, grad_req="write"):
        if self.params_initialized:
            arg_params, aux_params = self.get_params()
        if force_rebind:
            self.binded = False
            self._buckets = {}
        if self.binded:
            return
        self.for_training, self.inputs_need_grad, self._grad_req = for_training, inputs_need_grad, grad_req
        module = self._gen(self._default_bucket_key)
        module.bind(data_shapes, label_shapes, for_training, inputs_need_grad, force_rebind=False, shared_module=None,
                    grad_req=grad_req)
        self._curr_module = module
        self._curr_bucket_key = self._default_bucket_key
        self._buckets[self._default_bucket_key] = module
        self.binded = True
        if self.params_initialized:
            self.set_params(arg_params, aux_params)

    def switch_bucket(self, bucket_key, data_shapes, label_shapes=None):
        assert self.binded
        if bucket_key not in self._buckets:
            module = self._gen(bucket_key)
            module.bind(data_shapes, label_shapes, self._curr_module.for_training, self._curr_module.inputs_need_grad,
                        force_rebind=False, shared_module=self._buckets[self._default_bucket_key], grad_req=self._grad_req)
            self._buckets[bucket_key] = module
        self._curr_module = self._buckets[bucket_key]
        self._curr_bucket_key = bucket_key

    def init_optimizer(self, kvstore="local", optimizer="sgd", optimizer_params=(("learning_rate", 0.01),),
                       force_init=False, **kw):
        assert self.binded and self.params_initialized
        if self.optimizer_initialized and not force_init:
            return
        self._curr_module.init_optimizer(kvstore, optimizer, optimizer_params, force_init=force_init)
        for mod in self._buckets.values():
            if mod is not self._curr_module:
                mod.borrow_optimizer(self._curr_module)
        self.optimizer_initialized = True

    def prepare(self, data_batch, sparse_row_id_fn=None):
        original = self._curr_bucket_key
        self.switch_bucket(data_batch.bucket_key, data_batch.provide_data, data_batch.provide_label)
        self.switch_bucket(original, None, None)

    def forward(self, data_batch, is_train=None):
        self.switch_bucket(data_batch.bucket_key, data_batch.provide_data, data_batch.provide_label)
        self._curr_module.forward(data_batch, is_train=is_train)

    def backward(self, out_grads=None):
        self._curr_module.backward(out_grads=out_grads)

    def update(self):
        self._params_dirty = True
        self._curr_module.update()

    def get_outputs(self, merge_multi_context=True):
        return self._curr_module.get_outputs(merge_multi_context)

    def get_input_grads(self, merge_multi_context=True):
        return self._curr_module.get_input_grads(merge_multi_context)

    def update_metric(self, eval_metric, labels, pre_sliced=False):
        self._curr_module.update_metric(eval_metric, labels, pre_sliced)

    def install_monitor(self, mon):
        for mod in self._buckets.values():
            mod.install_monitor(mon)


class SequentialModule(BaseModule):
    """Chain of modules; outputs of one feed the next (sequential_module.py)."""
    META_TAKE_LABELS = "take_labels"
    META_AUTO_WIRING = "auto_wiring"

    def __init__(self, logger=logging):
        super(SequentialModule, self).__init__(logger=logger)
        self._modules, self._metas = [], []
        self._label_shapes = self._data_shapes = None

    def add(self, module, **kwargs):
        self._modules.append(module)
        self._metas.append(kwargs)
        self.binded = self.params_initialized = self.optimizer_initialized = False
        return self

    @property
    def data_names(self):
        return self._modules[0].data_names if self._modules else []

    @property
    def output_names(self):
        return self._modules[-1].output_names if self._modules else []

    @property
    def data_shapes(self):
        return self._modules[0].data_shapes

    @property
    def label_shapes(self):
        return self._label_shapes

    @property
    def output_shapes(self):
        return self._modules[-1].output_shapes

    def get_params(self):
        arg_params, aux_params = {}, {}
        for module in self._modules:
            arg, aux = module.get_params()
            arg_params.update(arg)
            aux_params.update(aux)
        return (arg_params, aux_params)

    def init_params(self, initializer=Uniform(0.01), arg_params=None, aux_params=None, allow_missing=False,
                    force_init=False, allow_extra=False):
        if self.params_initialized and not force_init:
            return
        for module in self._modules:
            module.init_params(initializer, arg_params, aux_params, allow_missing=True if arg_params else allow_missing,
                               force_init=force_init, allow_extra=True)
        self.params_initialized = True

    def bind(self, data_shapes, label_shapes=None, for_training=True, inputs_need_grad=False, force_rebind=False,
             shared_module=None, grad_req="write"):
        if self.binded and not force_rebind:
            return
        self.for_training, self.inputs_need_grad = for_training, inputs_need_grad
        self._label_shapes = label_shapes
        my_data_shapes = data_shapes
        anybody_ever_needs_label = False
        for i, (module, meta) in enumerate(zip(self._modules, self._metas)):
            my_label_shapes = label_shapes if meta.get(self.META_TAKE_LABELS, False) else None
            anybody_ever_needs_label |= my_label_shapes is not None
            my_inputs_need_grad = bool(inputs_need_grad or (for_training and i > 0))
            if meta.get(self.META_AUTO_WIRING, False):
                data_names = module.data_names
                my_data_shapes = [(new_name, shape) for (new_name, (_, shape)) in zip(data_names, [(d[0], d[1]) for d in my_data_shapes])]
            module.bind(my_data_shapes, my_label_shapes, for_training, my_inputs_need_grad, force_rebind, None, grad_req)
            # shapes for the next module need one forward to be known: use symbolic inference
            shapes = dict((d[0], d[1]) for d in my_data_shapes)
            _a, out_shapes, _x = module.symbol.infer_shape(**{k: v for k, v in shapes.items()})
            my_data_shapes = list(zip(module.output_names, out_shapes))
        if not anybody_ever_needs_label:
            self._label_shapes = None
        self.binded = True

    def init_optimizer(self, kvstore="local", optimizer="sgd", optimizer_params=(("learning_rate", 0.01),),
                       force_init=False, **kw):
        if self.optimizer_initialized and not force_init:
            return
        for module in self._modules:
            module.init_optimizer(kvstore, optimizer, optimizer_params, force_init=force_init)
        self.optimizer_initialized = True

    def forward(self, data_batch, is_train=None):
        from ..io import DataBatch
        batch = data_batch
        for i, module in enumerate(self._modules):
            module.forward(batch, is_train=is_train)
            if i + 1 == len(self._modules):
                break
            batch = DataBatch(module.get_outputs(), data_batch.label, pad=data_batch.pad)

    def backward(self, out_grads=None):
        for i, module in reversed(list(enumerate(self._modules))):
            module.backward(out_grads=out_grads)
            if i == 0:
                break
            out_grads = module.get_input_grads()

    def update(self):
        for module in self._modules:
            module.update()

    def get_outputs(self, merge_multi_context=True):
        return self._modules[-1].get_outputs(merge_multi_context)

    def get_input_grads(self, merge_multi_context=True):
        return self._modules[0].get_input_grads(merge_multi_context)

    def update_metric(self, eval_metric, labels, pre_sliced=False):
        for meta, module in zip(self._metas, self._modules):
            if meta.get(self.META_TAKE_LABELS, False):
                module.update_metric(eval_metric, labels, pre_sliced)

    def install_monitor(self, mon):
        for module in self._modules:
            module.install_monitor(mon)


class PythonModule(BaseModule):
    """Module whose computation is written in python (python_module.py): no parameters by default."""

    def __init__(self, data_names, label_names, output_names, logger=logging):
        super(PythonModule, self).__init__(logger=logger)
        self._data_names = list(data_names)
        self._label_names = list(label_names) if label_names is not None else []
        self._output_names = list(output_names)
        self._data_shapes = self._label_shapes = self._output_shapes = None

    @property
    def data_names(self):
        return self._data_names

    @property
    def output_names(self):
        return self._output_names

    @property
    def data_shapes(self):
        return self._data_shapes

    @property
    def label_shapes(self):
        return self._label_shapes

    @property
    def output_shapes(self):
        return self._output_shapes

    def get_params(self):
        return (dict(), dict())

    def init_params(self, *a, **k):
        self.params_initialized = True

    def update(self):
        pass

    def update_metric(self, eval_metric, labels, pre_sliced=False):
        if self._label_shapes is None:
            return
        eval_metric.update(labels, self.get_outputs())

    def bind(self, data_shapes, label_shapes=None, for_training=True, inputs_need_grad=False, force_rebind=False,
             shared_module=None, grad_req="write"):
        if self.binded and not force_rebind:
            return
        self.for_training, self.inputs_need_grad = for_training, inputs_need_grad
        self._data_shapes, self._label_shapes = data_shapes, label_shapes
        self._output_shapes = self._compute_output_shapes()
        self.binded = True

    def _compute_output_shapes(self):
        raise NotImplementedError()

    def init_optimizer(self, *a, **k):
        self.optimizer_initialized = True

    def install_monitor(self, mon):
        pass


class PythonLossModule(PythonModule):
    def __init__(self, name="pyloss", data_names=("data",), label_names=("softmax_label",), logger=logging, grad_func=None):
        super(PythonLossModule, self).__init__(data_names, label_names, [name + "_output"], logger=logger)
        self._name = name
        self._scores = self._labels = self._scores_grad = None
        self._grad_func = grad_func

    def _compute_output_shapes(self):
        return [(self._name + "_output", self._data_shapes[0][1])]

    def forward(self, data_batch, is_train=None):
        self._scores = data_batch.data[0]
        if is_train is None:
            is_train = self.for_training
        if is_train:
            self._labels = data_batch.label[0]

    def get_outputs(self, merge_multi_context=True):
        return [self._scores]

    def backward(self, out_grads=None):
        assert out_grads is None, "For a loss module, out_grads should be None"
        assert self.for_training
        if self._grad_func is not None:
            grad = self._grad_func(self._scores, self._labels)
            if not isinstance(grad, nd.NDArray):
                grad = nd.array(grad)
            self._scores_grad = grad
        else:
            raise NotImplementedError()

    def get_input_grads(self, merge_multi_context=True):
        return [self._scores_grad]
