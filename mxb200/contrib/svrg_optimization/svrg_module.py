"""SVRGModule: every ``update_freq`` epochs the full-batch gradient mu at a snapshot w~ is computed; each step then
uses  g_i(w) - g_i(w~) + mu  instead of g_i(w) (svrg_module.py:30-580).  Implemented with two Modules sharing the
symbol: ``self`` holds w, ``_mod_aux`` holds the snapshot w~."""
import logging
import time

from ... import metric as _metric
from ...module.module import Module
from ...ndarray import zeros_like
from .svrg_optimizer import _AssignmentOptimizer, _SVRGOptimizer  # noqa: F401  (reference module layout)


class SVRGModule(Module):
    def __init__(self, symbol, data_names=("data",), label_names=("softmax_label",), logger=logging, context=None,
                 work_load_list=None, fixed_param_names=None, state_names=None, group2ctxs=None, compression_params=None,
                 update_freq=None):
        super(SVRGModule, self).__init__(symbol, data_names=data_names, label_names=label_names, logger=logger, context=context,
                                         work_load_list=work_load_list, fixed_param_names=fixed_param_names,
                                         state_names=state_names, group2ctxs=group2ctxs, compression_params=compression_params)
        if not isinstance(update_freq, int) or update_freq <= 0:
            raise ValueError("update_freq in SVRGModule must be a positive integer to represent the frequency for "
                             "calculating full gradients")
        self.update_freq = update_freq
        self._mod_aux = Module(symbol, data_names, label_names, logger, context, work_load_list, fixed_param_names, state_names,
                               group2ctxs, compression_params)
        self._param_dict = None

    def bind(self, data_shapes, label_shapes=None, for_training=True, inputs_need_grad=False, force_rebind=False,
             shared_module=None, grad_req="write"):
        super(SVRGModule, self).bind(data_shapes, label_shapes, for_training, inputs_need_grad, force_rebind, shared_module, grad_req)
        if for_training:
            self._mod_aux.bind(data_shapes, label_shapes, for_training, inputs_need_grad, force_rebind, shared_module, grad_req)

    def init_params(self, *a, **kw):
        super(SVRGModule, self).init_params(*a, **kw)
        if self._mod_aux.binded:
            arg, aux = self.get_params()
            self._mod_aux.init_params(arg_params=arg, aux_params=aux, allow_missing=False, force_init=True)

    def init_optimizer(self, kvstore="local", optimizer="sgd", optimizer_params=(("learning_rate", 0.01),), force_init=False):
        """Install ``_SVRGOptimizer`` around the user's optimizer (svrg_module.py:114-151 of the reference): ordinary keys
        are updated by ``optimizer``; keys named ``<param>_full`` are ASSIGNED, which is how several workers accumulate the
        full gradients in a kvstore."""
        from ... import model as _model
        from ...ndarray import zeros
        batch_size = self._exec_group.batch_size
        arg_params = self.get_params()[0]
        kv, update_on_kvstore = _model._create_kvstore(kvstore, len(self._context), arg_params)
        if kv is not None and "dist" in kv.type and "_sync" in kv.type:
            batch_size *= kv.num_workers
        names = list(self._exec_group.param_names)
        nctx = len(self._context)
        idx2name = {}
        if update_on_kvstore:
            idx2name.update(enumerate(names))
        else:
            for k in range(nctx):
                idx2name.update({i * nctx + k: n for i, n in enumerate(names)})
        for n in names:
            idx2name[max(idx2name) + 1] = n + "_full"
        params = dict(optimizer_params)
        params.setdefault("rescale_grad", 1.0 / batch_size)
        params["param_idx2name"] = idx2name
        svrg_opt = _SVRGOptimizer(default_optimizer=optimizer, **params)
        super(SVRGModule, self).init_optimizer(kvstore=kv if kv is not None else None, optimizer=svrg_opt,
                                               optimizer_params=optimizer_params, force_init=force_init)
        if self._kvstore is not None:
            for idx, n in enumerate(names):
                self._kvstore.init(n + "_full", zeros(arg_params[n].shape))

    def forward(self, data_batch, is_train=None):
        super(SVRGModule, self).forward(data_batch, is_train)
        if is_train is None or is_train:
            if self._mod_aux.binded:
                self._mod_aux.forward(data_batch, is_train=True)

    def backward(self, out_grads=None):
        super(SVRGModule, self).backward(out_grads)
        if self._mod_aux.binded:
            self._mod_aux.backward(out_grads)

    def update_full_grads(self, train_data):
        """Snapshot w~ <- w and mu <- mean over the data set of grad(w~)."""
        arg, aux = self.get_params()
        self._mod_aux.set_params(arg_params=arg, aux_params=aux)
        train_data.reset()
        nbatch = 0
        acc = None
        for batch in train_data:
            self._mod_aux.forward(batch, is_train=True)
            self._mod_aux.backward()
            grads = self._grad_dict(self._mod_aux)
            if acc is None:
                acc = {k: zeros_like(v) for k, v in grads.items()}
            for k, v in grads.items():
                acc[k] += v
            nbatch += 1
        self._param_dict = {k: v / max(nbatch, 1) for k, v in (acc or {}).items()}
        train_data.reset()

    @staticmethod
    def _grad_dict(mod):
        names = mod._exec_group.param_names
        out = {}
        for name, grads in zip(names, mod._exec_group.grad_arrays):
            if grads and grads[0] is not None:
                g = grads[0].copy()
                for other in grads[1:]:
                    g += other.as_in_context(g.context)
                out[name] = g
        return out

    def _update_svrg_gradients(self):
        """g <- g(w) - g(w~) + mu, written into the main module's gradient arrays before the optimizer runs."""
        if self._param_dict is None:
            return
        names = self._exec_group.param_names
        for name, g_cur, g_snap in zip(names, self._exec_group.grad_arrays, self._mod_aux._exec_group.grad_arrays):
            if not g_cur or g_cur[0] is None or name not in self._param_dict:
                continue
            for d, (gc, gs) in enumerate(zip(g_cur, g_snap)):
                gc[:] = gc - gs + self._param_dict[name].as_in_context(gc.context)

    def update(self):
        self._update_svrg_gradients()
        super(SVRGModule, self).update()

    def fit(self, train_data, eval_data=None, eval_metric="acc", epoch_end_callback=None, batch_end_callback=None,
            kvstore="local", optimizer="sgd", optimizer_params=(("learning_rate", 0.01),), eval_end_callback=None,
            eval_batch_end_callback=None, initializer=None, arg_params=None, aux_params=None, allow_missing=False,
            force_rebind=False, force_init=False, begin_epoch=0, num_epoch=None, validation_metric=None, monitor=None,
            sparse_row_id_fn=None):
        assert num_epoch is not None, "please specify number of epochs"
        from ... import initializer as _init
        from ...model import BatchEndParam
        self.bind(data_shapes=train_data.provide_data, label_shapes=train_data.provide_label, for_training=True, force_rebind=force_rebind)
        self.init_params(initializer=initializer or _init.Uniform(0.01), arg_params=arg_params, aux_params=aux_params,
                         allow_missing=allow_missing, force_init=force_init)
        self.init_optimizer(kvstore=kvstore, optimizer=optimizer, optimizer_params=optimizer_params)
        if not isinstance(eval_metric, _metric.EvalMetric):
            eval_metric = _metric.create(eval_metric)
        validation_metric = validation_metric or eval_metric
        for epoch in range(begin_epoch, num_epoch):
            tic = time.time()
            eval_metric.reset()
            if epoch % self.update_freq == 0:
                self.update_full_grads(train_data)
            for nbatch, batch in enumerate(train_data):
                self.forward_backward(batch)
                self.update()
                self.update_metric(eval_metric, batch.label)
                if batch_end_callback is not None:
                    p = BatchEndParam(epoch=epoch, nbatch=nbatch, eval_metric=eval_metric, locals=locals())
                    for cb in (batch_end_callback if isinstance(batch_end_callback, list) else [batch_end_callback]):
                        cb(p)
            for name, val in eval_metric.get_name_value():
                self.logger.info("Epoch[%d] Train-%s=%f", epoch, name, val)
            self.logger.info("Epoch[%d] Time cost=%.3f", epoch, time.time() - tic)
            arg, aux = self.get_params()
            if epoch_end_callback is not None:
                for cb in (epoch_end_callback if isinstance(epoch_end_callback, list) else [epoch_end_callback]):
                    cb(epoch, self.symbol, arg, aux)
            if eval_data is not None:
                res = self.score(eval_data, validation_metric, score_end_callback=eval_end_callback,
                                 batch_end_callback=eval_batch_end_callback, epoch=epoch)
                for name, val in res:
                    self.logger.info("Epoch[%d] Validation-%s=%f", epoch, name, val)
            train_data.reset()
