"""Gluon Parameter / ParameterDict (reference: python/mxnet/gluon/parameter.py).

A Parameter keeps one data array and one gradient array per context (``_data`` / ``_grad`` lists,
parameter.py:496-540).  ``reset_ctx`` = reduce (mean of the per-ctx copies) + re-initialise on the new
context list (parameter.py:312-323,393-416): the single-box analogue of an elastic re-init (SURVEY.md
Appendix D).

Provenance: adapted from Apache MXNet (incubating) 1.3.1 ``python/mxnet/gluon/parameter.py`` (Apache License 2.0).  The public API, argument
names and control flow follow the original so that user scripts keep working; the tensor maths runs on torch /
the mxb200 kernels.  This file is API-compatibility glue, not original B200 work ("Provenance").
"""
import re
import warnings
from collections import OrderedDict

import numpy as np
import torch

from ..base import MXNetError, torch_dtype, np_dtype
from ..context import Context, cpu, current_context
from .. import initializer, autograd
from ..ndarray import NDArray, zeros, sparse as _sp
from ..ndarray import serialization as _ser

__all__ = ["DeferredInitializationError", "Parameter", "Constant", "ParameterDict", "tensor_types"]

tensor_types = (NDArray,)


class DeferredInitializationError(MXNetError):
    """Error for unfinished deferred initialization."""


class Parameter(object):
    def __init__(self, name, grad_req="write", shape=None, dtype="float32", lr_mult=1.0, wd_mult=1.0, init=None,
                 allow_deferred_init=False, differentiable=True, stype="default", grad_stype="default"):
        self._var = None
        self._data = None
        self._grad = None
        self._ctx_list = None
        self._ctx_map = None
        self._trainer = None
        self._deferred_init = ()
        self._differentiable = differentiable
        self._allow_deferred_init = allow_deferred_init
        self._grad_req = None
        if isinstance(shape, int):
            shape = (shape,)
        self._shape = tuple(shape) if shape is not None else None
        self.name = name
        self.dtype = dtype
        self.lr_mult = lr_mult
        self.wd_mult = wd_mult
        self.grad_req = grad_req
        self.init = init
        valid = ("default", "row_sparse", "csr")
        assert grad_stype in valid, "grad_stype for Parameter '%s' must be one of 'default', 'row_sparse', or 'csr', " \
                                    "but got '%s'" % (name, grad_stype)
        assert stype in valid, "stype for Parameter '%s' must be one of 'default', 'row_sparse', or 'csr', but got '%s'" \
                               % (name, stype)
        self._stype = stype
        self._grad_stype = grad_stype
        self._fresh_grad = False
        self._grad_sink = None     # fp32 gradient-arena view when a fused Trainer owns this parameter

    def __repr__(self):
        return "Parameter {name} (shape={shape}, dtype={dtype})".format(name=self.name, shape=self.shape, dtype=self.dtype)

    @property
    def grad_req(self):
        return self._grad_req

    @grad_req.setter
    def grad_req(self, req):
        assert req in ["write", "add", "null"], "grad_req must be one of 'write', 'add', or 'null', but got '%s'" % req
        if not self._differentiable:
            req = "null"
        if self._grad_req == req:
            return
        self._grad_req = req
        if req == "null" and self._grad is not None:
            self._grad = None
            for d in self._data or []:
                d._t.requires_grad_(False)
        elif self._data is not None:
            self._init_grad()

    @property
    def shape(self):
        return self._shape

    @shape.setter
    def shape(self, new_shape):
        if self._shape is None:
            self._shape = tuple(new_shape)
            return
        assert len(self._shape) == len(new_shape) and all(j in (0, i) for i, j in zip(new_shape, self._shape)), \
            "Expected shape %s is incompatible with given shape %s." % (str(new_shape), str(self._shape))
        self._shape = tuple(new_shape)

    def _check_and_get(self, arr_list, ctx):
        if arr_list is not None:
            if ctx is list:
                return arr_list
            if ctx is None:
                if len(arr_list) == 1:
                    return arr_list[0]
                ctx = current_context()
            for c, a in zip(self._ctx_list, arr_list):
                if c == ctx:
                    return a
            raise RuntimeError("Parameter '%s' was not initialized on context %s. It was only initialized on %s."
                               % (self.name, str(ctx), str(self._ctx_list)))
        if self._deferred_init:
            raise DeferredInitializationError(
                "Parameter '%s' has not been initialized yet because initialization was deferred. Actual "
                "initialization happens during the first forward pass. Please pass one batch of data through the "
                "network before accessing Parameters." % self.name)
        raise RuntimeError("Parameter '%s' has not been initialized. Note that you should initialize parameters and "
                           "create Trainer with Block.collect_params() instead of Block.params because the later "
                           "does not include Parameters of nested child Blocks" % self.name)

    def _load_init(self, data, ctx):
        if self.shape:
            for self_dim, data_dim in zip(self.shape, data.shape):
                assert self_dim in (0, data_dim), "Failed loading Parameter '%s' from saved params: shape incompatible " \
                    "expected %s vs saved %s" % (self.name, str(self.shape), str(data.shape))
            self.shape = tuple(i if i != 0 else j for i, j in zip(self.shape, data.shape))
        if self.dtype:
            if torch_dtype(self.dtype) != data._t.dtype:
                data = data.astype(self.dtype)
        if isinstance(ctx, Context):
            ctx = [ctx]
        if self._data is None:
            if self._deferred_init:
                assert ctx is None or set(ctx) == set(self._deferred_init[1]), \
                    "Failed to load Parameter '%s' on %s because it was previous initialized on %s." % (
                        self.name, str(ctx), str(self.list_ctx()))
                ctx = self._deferred_init[1]
            elif ctx is None:
                ctx = [cpu()]
            self._init_impl(data, ctx)
        else:
            assert ctx is None or set(ctx) == set(self.list_ctx()), \
                "Failed to load Parameter '%s' on %s because it was previous initialized on %s." % (
                    self.name, str(ctx), str(self.list_ctx()))
            self.set_data(data)
        self._deferred_init = ()

    def _finish_deferred_init(self):
        if not self._deferred_init:
            return
        init, ctx, default_init, data = self._deferred_init
        self._deferred_init = ()
        assert self.shape is not None and np.prod(self.shape) > 0, \
            "Cannot initialize Parameter '%s' because it has invalid shape: %s. Please specify in_units, " \
            "in_channels, etc for `Block`s." % (self.name, str(self.shape))
        with autograd.pause():
            if data is None:
                data = zeros(self.shape, dtype=self.dtype, ctx=cpu())
                initializer.create(default_init)(
                    initializer.InitDesc(self.name, {"__init__": init.dumps() if isinstance(init, initializer.Initializer) else (init or "")}),
                    data)
            self._init_impl(data, ctx)

    def _init_impl(self, data, ctx_list):
        self._ctx_list = list(ctx_list)
        self._ctx_map = {}
        self._data = [data.copyto(ctx) for ctx in self._ctx_list]
        self._init_grad()

    def _init_grad(self):
        if self.grad_req == "null":
            self._grad = None
            return
        self._grad = []
        for d in self._data:
            d.attach_grad(self.grad_req)
            self._grad.append(NDArray(d._t.grad, d.context))

    def _reduce(self):
        """Average of the per-context copies, on CPU (parameter.py:312-323)."""
        block = self._raw_list_data()
        acc = block[0]._t.detach().to("cpu").float()
        for b in block[1:]:
            acc = acc + b._t.detach().to("cpu").float()
        return NDArray((acc / len(block)).to(block[0]._t.dtype), cpu())

    def initialize(self, init=None, ctx=None, default_init=initializer.Uniform(), force_reinit=False):
        if self._data is not None and not force_reinit:
            warnings.warn("Parameter '%s' is already initialized, ignoring. Set force_reinit=True to re-initialize."
                          % self.name, stacklevel=2)
            return
        if self._data is not None and self._trainer is not None and getattr(self._trainer, "_fused", None) is not None:
            self._trainer._reset_kvstore()        # the arena views are about to be replaced: un-fuse, re-fuse lazily
        self._data = self._grad = None
        if ctx is None:
            ctx = [current_context()]
        if isinstance(ctx, Context):
            ctx = [ctx]
        if init is None:
            init = default_init if self.init is None else self.init
        if not self.shape or np.prod(self.shape) <= 0:
            if self._allow_deferred_init:
                self._deferred_init = (init, ctx, default_init, None)
                return
            raise ValueError("Cannot initialize Parameter '%s' because it has invalid shape: %s." % (self.name, str(self.shape)))
        self._deferred_init = (init, ctx, default_init, None)
        self._finish_deferred_init()

    def reset_ctx(self, ctx):
        """Re-assign the Parameter to other contexts: mean of the old copies becomes the new value."""
        if ctx is None:
            ctx = [current_context()]
        if isinstance(ctx, Context):
            ctx = [ctx]
        if self._data:
            data = self._reduce()
            with autograd.pause():
                self._init_impl(data, ctx)
        elif self._deferred_init:
            init, _, default_init, data = self._deferred_init
            self._deferred_init = (init, ctx, default_init, data)
        else:
            raise ValueError("Cannot reset context for Parameter '%s' because it has not been initialized." % self.name)
        if self._trainer is not None:
            self._trainer._reset_kvstore()

    def set_data(self, data):
        self.shape = data.shape
        if self._data is None:
            assert self._deferred_init, "Parameter '%s' has not been initialized" % self.name
            self._deferred_init = self._deferred_init[:3] + (data,)
            return
        if self._trainer is not None and self._trainer._kv_initialized and self._trainer._update_on_kvstore:
            self._trainer._reset_kvstore()
        for arr in self._data:
            with torch.no_grad():
                arr._t.copy_(data._t.to(arr._t.device))
        self._refresh_fused_master()

    def _refresh_fused_master(self):
        """The low-precision weights were overwritten from outside (set_data / load_parameters): the fused Trainer must
        re-derive its fp32 master copy from them, otherwise its next step would resurrect the old values."""
        tr = self._trainer
        if tr is not None and getattr(tr, "_fused", None) is not None:
            tr._fused.refresh_master()

    def _check_row_sparse_access(self):
        if self._stype != "row_sparse":
            raise RuntimeError("Cannot return a copy of Parameter %s via row_sparse_data() because its storage type is %s. "
                               "Please use data() instead." % (self.name, self._stype))
        if getattr(self, "_trainer", None) is None:
            raise RuntimeError("Cannot get row_sparse data for Parameter '%s' when no Trainer is created with it." % self.name)

    def row_sparse_data(self, row_id):
        """Rows ``row_id`` of a row_sparse Parameter on ``row_id``'s context (reference: the Trainer pulls them from the
        kvstore, gluon/parameter.py:441-459)."""
        self._check_row_sparse_access()
        return self._rows_of(self._raw_data(row_id.context), row_id)

    def list_row_sparse_data(self, row_id):
        self._check_row_sparse_access()
        return [self._rows_of(d, row_id.as_in_context(d.context)) for d in self._raw_list_data()]

    @staticmethod
    def _rows_of(d, row_id):
        """Row-sparse view of the requested rows.  Its dense leaf is the parameter's own array, so an operator that consumes
        it under ``autograd.record()`` (Embedding, sparse dot) back-propagates into the parameter."""
        rsp = _sp.retain(_sp.cast_storage(d, "row_sparse"), row_id)
        if d._t.requires_grad:
            rsp.__dict__["_dense_leaf"] = d
        return rsp

    def _raw_data(self, ctx=None):
        """Dense backing array regardless of the declared storage type (framework-internal access)."""
        return self._check_and_get(self._data, ctx)

    def _raw_list_data(self):
        return self._check_and_get(self._data, list)

    def _check_dense_access(self, what):
        if self._stype != "default":
            raise RuntimeError("Cannot return a copy of Parameter '%s' via %s() because its storage type is %s. Please use "
                               "row_sparse_data() instead." % (self.name, what, self._stype))

    def data(self, ctx=None):
        self._check_dense_access("data")
        return self._check_and_get(self._data, ctx)

    def list_data(self):
        self._check_dense_access("list_data")
        return self._check_and_get(self._data, list)

    def grad(self, ctx=None):
        if self._data is not None and self._grad is None:
            raise RuntimeError("Cannot get gradient array for Parameter '%s' because grad_req='null'" % self.name)
        self._sync_grad()
        return self._grad_view(self._check_and_get(self._grad, ctx))

    def list_grad(self):
        if self._data is not None and self._grad is None:
            raise RuntimeError("Cannot get gradient array for Parameter '%s' because grad_req='null'" % self.name)
        self._sync_grad()
        return [self._grad_view(g) for g in self._check_and_get(self._grad, list)]

    def _raw_list_grad(self):
        """The dense gradient buffers themselves (framework-internal: kvstore push / pull targets)."""
        self._sync_grad()
        return self._check_and_get(self._grad, list)

    def _grad_view(self, g):
        """``grad_stype='row_sparse'`` (e.g. ``Embedding(sparse_grad=True)``): the gradient buffer is dense-backed (autograd
        accumulates into it) and is exposed with the declared storage type, like the reference's sparse gradient arrays."""
        st = getattr(self, "_grad_stype", "default")
        if st == "default" or g is None:
            return g
        from ..ndarray.sparse import BaseSparseNDArray, cast_storage
        return g if isinstance(g, BaseSparseNDArray) else cast_storage(g, st)

    def _sync_grad(self):
        # torch may replace .grad; keep our NDArray handles pointing at the live tensor
        if self._grad is None or self._data is None or self._grad_sink is not None:
            return
        for i, d in enumerate(self._data):
            g = d._t.grad
            if g is None:
                d._t.grad = self._grad[i]._t
            elif g is not self._grad[i]._t:
                self._grad[i]._t = g

    def list_ctx(self):
        if self._data is None:
            if self._deferred_init:
                return self._deferred_init[1]
            raise RuntimeError("Parameter '%s' has not been initialized" % self.name)
        return self._ctx_list

    def _set_trainer(self, trainer):
        """A row_sparse parameter can only be managed by one Trainer (parameter.py:188-199 of the reference)."""
        if getattr(self, "_stype", "default") != "default" and getattr(self, "_trainer", None) and trainer \
                and self._trainer is not trainer:
            raise RuntimeError("Failed to set the trainer for Parameter '%s' because it was already set. More than one "
                               "trainers for a %s Parameter is not supported." % (self.name, self._stype))
        self._trainer = trainer

    def zero_grad(self):
        if self._grad is None:
            return
        self._sync_grad()
        for g in self._grad:
            g._t.zero_()

    def var(self):
        if self._var is None:
            from .. import symbol
            self._var = symbol.var(self.name, shape=self.shape, dtype=self.dtype, lr_mult=self.lr_mult,
                                   wd_mult=self.wd_mult, init=self.init, stype=self._stype)
        return self._var

    def cast(self, dtype):
        self.dtype = dtype
        if self._data is None:
            return
        if self._trainer is not None and getattr(self._trainer, "_fused", None) is not None:
            self._trainer._reset_kvstore()        # parameters leave the arena: the Trainer re-fuses over the new tensors
        with autograd.pause():
            self._data = [NDArray(d._t.detach().to(torch_dtype(dtype)), d.context) for d in self._data]
            self._init_grad()


class Constant(Parameter):
    """A constant (non-trainable) parameter."""

    def __init__(self, name, value):
        if not isinstance(value, NDArray):
            from ..ndarray import array
            value = array(value)
        self.value = value

        class Init(initializer.Initializer):
            def _init_weight(self2, _, arr):
                arr[:] = value

            def dumps(self2):
                import json
                return json.dumps([init_name.lower(), {}])
        init_name = "Constant_{}_{}".format(name, id(self))
        Init.__name__ = init_name
        initializer._REG[init_name.lower()] = Init
        super(Constant, self).__init__(name, grad_req="null", shape=value.shape, dtype=value.dtype, init=init_name)


class ParameterDict(object):
    """Dictionary of Parameters with a shared prefix (parameter.py:560-920)."""

    def __init__(self, prefix="", shared=None):
        self._prefix = prefix
        self._params = OrderedDict()
        self._shared = shared

    def __repr__(self):
        s = "{name}(\n{content}\n)"
        name = self._prefix + " " if self._prefix else ""
        return s.format(name=name, content="\n".join(["  " + repr(v) for v in self.values()]))

    def __getitem__(self, key):
        return self._params[key]

    def __iter__(self):
        return iter(self._params)

    def __len__(self):
        return len(self._params)

    def __contains__(self, key):
        return key in self._params

    def items(self):
        return self._params.items()

    def keys(self):
        return self._params.keys()

    def values(self):
        return self._params.values()

    @property
    def prefix(self):
        return self._prefix

    def _get_impl(self, name):
        if name in self._params:
            return self._params[name]
        if self._shared is not None and name in self._shared._params:
            self._params[name] = self._shared._params[name]
            return self._shared._params[name]
        return None

    def get(self, name, **kwargs):
        name = self.prefix + name
        param = self._get_impl(name)
        if param is None:
            param = Parameter(name, **kwargs)
            self._params[name] = param
        else:
            for k, v in kwargs.items():
                if hasattr(param, k) and getattr(param, k) is not None:
                    existing = getattr(param, k)
                    if k == "shape" and v is not None and len(v) == len(existing):
                        inferred, matched = [], True
                        for d1, d2 in zip(v, existing):
                            if d1 != d2 and d1 * d2 != 0:
                                matched = False
                                break
                            inferred.append(d1 if d1 != 0 else d2)
                        if matched:
                            param._shape = tuple(inferred)
                            continue
                    assert v is None or v == existing or (k == "dtype" and np_dtype(torch_dtype(v)) == np_dtype(torch_dtype(existing))), \
                        "Cannot retrieve Parameter '%s' because desired attribute does not match with stored for " \
                        "attribute '%s': desired '%s' vs stored '%s'." % (name, k, str(v), str(getattr(param, k)))
                else:
                    setattr(param, k, v)
        return param

    def get_constant(self, name, value=None):
        name = self.prefix + name
        param = self._get_impl(name)
        if param is None:
            if value is None:
                raise KeyError("No constant named '{}'. Please specify value if you want to create a new constant.".format(name))
            param = Constant(name, value)
            self._params[name] = param
        return param

    def update(self, other):
        for k, v in other.items():
            if k in self._params:
                assert self._params[k] is v, "Cannot update self with other because they have different Parameters " \
                                             "with the same name '%s'" % k
        for k, v in other.items():
            self._params[k] = v

    def initialize(self, init=initializer.Uniform(), ctx=None, verbose=False, force_reinit=False):
        if verbose:
            init.set_verbosity(verbose=verbose)
        for _, v in self.items():
            v.initialize(None, ctx, init, force_reinit=force_reinit)

    def zero_grad(self):
        for i in self.values():
            i.zero_grad()

    def reset_ctx(self, ctx):
        for i in self.values():
            i.reset_ctx(ctx)

    def setattr(self, name, value):
        for i in self.values():
            setattr(i, name, value)

    def save(self, filename, strip_prefix=""):
        arg_dict = {}
        for param in self.values():
            weight = param._reduce()
            if not param.name.startswith(strip_prefix):
                raise ValueError("Prefix '%s' is to be striped before saving, but Parameter's name '%s' does not start "
                                 "with '%s'." % (strip_prefix, param.name, strip_prefix))
            arg_dict[param.name[len(strip_prefix):]] = weight
        _ser.save(filename, arg_dict)

    def load(self, filename, ctx=None, allow_missing=False, ignore_extra=False, restore_prefix=""):
        if restore_prefix:
            for name in self.keys():
                assert name.startswith(restore_prefix), "restore_prefix is '%s' but Parameters name '%s' does not " \
                    "start with '%s'" % (restore_prefix, name, restore_prefix)
        lprefix = len(restore_prefix)
        loaded = [(k[4:] if k.startswith("arg:") or k.startswith("aux:") else k, v) for k, v in _ser.load(filename).items()]
        arg_dict = {restore_prefix + k: v for k, v in loaded}
        if not allow_missing:
            for name in self.keys():
                assert name in arg_dict, "Parameter '%s' is missing in file '%s', which contains parameters: %s." % (
                    name[lprefix:], filename, ", ".join(sorted(k[lprefix:] for k in arg_dict)[:8]))
        for name in arg_dict:
            if name not in self._params:
                assert ignore_extra, "Parameter '%s' loaded from file '%s' is not present in ParameterDict" % (
                    name[lprefix:], filename)
                continue
            self[name]._load_init(arg_dict[name], ctx)
