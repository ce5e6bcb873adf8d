"""Gluon Block / HybridBlock / SymbolBlock (reference: python/mxnet/gluon/block.py).

``hybridize()`` in the reference swaps the imperative execution for a CachedOp running a static graph
(src/imperative/cached_op.h:72).  Here hybridization selects the block's *fused* implementation when it
has one (``_fused_forward``: straight torch-level calls into our kernels with no per-op NDArray wrapping)
-- the same idea (drop the per-op frontend overhead, keep semantics) without a tracing compiler.
``export()`` traces ``hybrid_forward`` with ``F = mx.sym`` to write ``-symbol.json`` + ``-0000.params``.

Provenance: adapted from Apache MXNet (incubating) 1.3.1 ``python/mxnet/gluon/block.py`` (Apache License 2.0).  The public API, argument
names and control flow follow the original so that user scripts keep working; the tensor maths runs on torch /
the mxb200 kernels.  This file is API-compatibility glue, not original B200 work ("Provenance").
"""
import copy
import re
import threading
import warnings
from collections import OrderedDict

import numpy as np

from ..base import MXNetError
from ..context import cpu, current_context
from .. import ndarray as nd
from ..ndarray import NDArray
from ..ndarray import serialization as _ser
from .parameter import Parameter, ParameterDict, DeferredInitializationError
from .utils import _indent

__all__ = ["Block", "HybridBlock", "SymbolBlock"]


class _BlockScope(object):
    _current = threading.local()

    def __init__(self, block):
        self._block = block
        self._counter = {}
        self._old_scope = None
        self._name_scope = None

    @staticmethod
    def create(prefix, params, hint):
        current = getattr(_BlockScope._current, "value", None)
        if current is None:
            if prefix is None:
                if not hasattr(_BlockScope._current, "global_counter"):
                    _BlockScope._current.global_counter = {}
                cnt = _BlockScope._current.global_counter
                count = cnt.get(hint, 0)
                cnt[hint] = count + 1
                prefix = "%s%d_" % (hint, count)
            if params is None:
                params = ParameterDict(prefix)
            else:
                params = ParameterDict(params.prefix, params)
            return prefix, params
        if prefix is None:
            count = current._counter.get(hint, 0)
            prefix = "%s%d_" % (hint, count)
            current._counter[hint] = count + 1
        if params is None:
            parent = current._block.params
            params = ParameterDict(parent.prefix + prefix, parent._shared)
        else:
            params = ParameterDict(params.prefix, params)
        return current._block.prefix + prefix, params

    def __enter__(self):
        if self._block._empty_prefix:
            return self
        self._old_scope = getattr(_BlockScope._current, "value", None)
        _BlockScope._current.value = self
        from .. import name as _name
        self._name_scope = _name.Prefix(self._block.prefix)     # symbols created inside are named <block prefix><hint><n>
        self._name_scope.__enter__()
        return self

    def __exit__(self, ptype, value, trace):
        if self._block._empty_prefix:
            return
        self._name_scope.__exit__(ptype, value, trace)
        self._name_scope = None
        _BlockScope._current.value = self._old_scope


def _flatten(args, inout_str):
    if isinstance(args, NDArray):
        return [args], int(0)
    if args is None:
        return [None], int(-1)
    from ..symbol import Symbol
    if isinstance(args, Symbol):
        length = len(args.list_outputs())
        length = length if length > 1 else 0
        return [args], int(length)
    assert isinstance(args, (list, tuple)), "HybridBlock %s must be (nested) list of Symbol or NDArray, but got %s " \
                                            "of type %s" % (inout_str, str(args), str(type(args)))
    flat, fmts = [], []
    for i in args:
        arg, fmt = _flatten(i, inout_str)
        flat.extend(arg)
        fmts.append(fmt)
    return flat, fmts


def _regroup(args, fmt):
    if isinstance(fmt, int):
        if fmt <= 0:
            return args[0], args[1:]
        return args[:fmt], args[fmt:]
    ret = []
    for i in fmt:
        res, args = _regroup(args, i)
        ret.append(res)
    return ret, args


class Block(object):
    def __init__(self, prefix=None, params=None):
        self._empty_prefix = prefix == ""
        self._prefix, self._params = _BlockScope.create(prefix, params, self._alias())
        self._name = self._prefix[:-1] if self._prefix.endswith("_") else self._prefix
        self._scope = _BlockScope(self)
        self._children = OrderedDict()
        self._reg_params = {}
        self._forward_hooks = OrderedDict()
        self._forward_pre_hooks = OrderedDict()

    def __repr__(self):
        s = "{name}(\n{modstr}\n)"
        modstr = "\n".join(["  ({key}): {block}".format(key=key, block=_indent(block.__repr__(), 2))
                            for key, block in self.__dict__.items() if isinstance(block, Block)])
        return s.format(name=self.__class__.__name__, modstr=modstr)

    def __setattr__(self, name, value):
        if hasattr(self, name):
            existing = getattr(self, name)
            if isinstance(existing, (Parameter, Block)) and not isinstance(value, type(existing)):
                raise TypeError("Changing attribute type for {name} from {type1} to {type2} is not allowed.".format(
                    name=name, type1=type(existing), type2=type(value)))
        if isinstance(value, Block):
            self.register_child(value, name)
        elif isinstance(value, Parameter):
            assert name not in self._reg_params, "Overriding Parameter attribute %s is not allowed. " \
                "If you want to share parameters between blocks, please set 'params' at Block construction instead."
            self._reg_params[name] = value
        super(Block, self).__setattr__(name, value)

    def _check_container_with_block(self):
        children = set(self._children.values())

        def _find(data):
            if isinstance(data, (list, tuple)):
                return any(_find(e) for e in data)
            if isinstance(data, dict):
                return any(_find(v) for v in data.values())
            return isinstance(data, Block) and data not in children
        for k, v in self.__dict__.items():
            if isinstance(v, (list, tuple, dict)) and not (k.startswith("__") or k == "_children"):
                if _find(v):
                    warnings.warn('"{name}" is an unregistered container with Blocks. Note that Blocks inside the '
                                  "list, tuple or dict will not be registered automatically. Make sure to register "
                                  "them using register_child() or switching to nn.Sequential/nn.HybridSequential "
                                  "instead. ".format(name=self.__class__.__name__ + "." + k), stacklevel=3)

    def _alias(self):
        return self.__class__.__name__.lower()

    @property
    def prefix(self):
        return self._prefix

    @property
    def name(self):
        return self._name

    def name_scope(self):
        return self._scope

    @property
    def params(self):
        return self._params

    def collect_params(self, select=None):
        self._check_container_with_block()
        ret = ParameterDict(self._params.prefix)
        if not select:
            ret.update(self.params)
        else:
            pattern = re.compile(select)
            ret.update({name: value for name, value in self.params.items() if pattern.match(name)})
        for cld in self._children.values():
            ret.update(cld.collect_params(select=select))
        return ret

    def _collect_params_with_prefix(self, prefix=""):
        if prefix:
            prefix += "."
        ret = {prefix + key: val for key, val in self._reg_params.items()}
        for name, child in self._children.items():
            ret.update(child._collect_params_with_prefix(prefix + name))
        return ret

    def save_parameters(self, filename):
        """Structural names without prefix (block.py:315-335)."""
        params = self._collect_params_with_prefix()
        arg_dict = {key: val._reduce() for key, val in params.items()}
        _ser.save(filename, arg_dict)

    def save_params(self, filename):
        warnings.warn("save_params is deprecated. Please use save_parameters.")
        try:
            self.collect_params().save(filename, strip_prefix=self.prefix)
        except ValueError as e:
            raise ValueError("%s\nsave_params is deprecated. Using save_parameters may resolve this error." % e)

    def load_parameters(self, filename, ctx=None, allow_missing=False, ignore_extra=False):
        loaded = _ser.load(filename)
        params = self._collect_params_with_prefix()
        if not loaded and not params:
            return
        if not any("." in i for i in loaded.keys()):
            # legacy loading
            del loaded
            self.collect_params().load(filename, ctx, allow_missing, ignore_extra, self.prefix)
            return
        if not allow_missing:
            for name in params.keys():
                assert name in loaded, "Parameter '%s' is missing in file '%s', which contains parameters: %s. " \
                    "Set allow_missing=True to ignore missing parameters." % (name, filename, ", ".join(sorted(loaded)[:8]))
        for name in loaded:
            if not ignore_extra and name not in params:
                raise ValueError("Parameter '%s' loaded from file '%s' is not present in ParameterDict, which "
                                 "contains parameters %s. Set ignore_extra=True to ignore. " % (
                                     name, filename, ", ".join(sorted(params)[:8])))
            if name in params:
                params[name]._load_init(loaded[name], ctx)

    def load_params(self, filename, ctx=None, allow_missing=False, ignore_extra=False):
        warnings.warn("load_params is deprecated. Please use load_parameters.")
        self.load_parameters(filename, ctx, allow_missing, ignore_extra)

    def register_child(self, block, name=None):
        if name is None:
            name = str(len(self._children))
        self._children[name] = block

    def register_forward_pre_hook(self, hook):
        """hook(block, inputs) runs before forward; returns a handle whose ``detach()`` removes it (block.py:356-386)."""
        handle = HookHandle()
        handle.attach(self._forward_pre_hooks, hook)
        return handle

    def register_forward_hook(self, hook):
        """hook(block, inputs, outputs) runs after forward; returns a detachable handle."""
        handle = HookHandle()
        handle.attach(self._forward_hooks, hook)
        return handle

    def apply(self, fn):
        for cld in self._children.values():
            cld.apply(fn)
        fn(self)
        return self

    def initialize(self, init=None, ctx=None, verbose=False, force_reinit=False):
        from .. import initializer
        self.collect_params().initialize(init if init is not None else initializer.Uniform(), ctx, verbose, force_reinit)

    def hybridize(self, active=True, **kwargs):
        for cld in self._children.values():
            cld.hybridize(active, **kwargs)

    def cast(self, dtype):
        for child in self._children.values():
            child.cast(dtype)
        for _, param in self.params.items():
            param.cast(dtype)

    def __call__(self, *args):
        for hook in self._forward_pre_hooks.values():
            hook(self, args)
        out = self.forward(*args)
        for hook in self._forward_hooks.values():
            hook(self, args, out)
        return out

    def forward(self, *args):
        raise NotImplementedError

    def summary(self, *inputs):
        summary = OrderedDict()
        hooks = []
        assert not getattr(self, "_active", False), "\"{}\" must not be hybridized to print summary.".format(self.name)

        def _register(block):
            def hook(blk, _inp, outputs):
                idx = len(summary)
                key = "%s-%i" % (blk.__class__.__name__, idx + 1)
                outs = outputs if isinstance(outputs, (list, tuple)) else [outputs]
                nparams = sum(int(np.prod(p.shape)) for p in blk.params.values() if p.shape)
                summary[key] = {"output_shape": [o.shape for o in outs if hasattr(o, "shape")], "n_params": nparams}
            hooks.append((block, block.register_forward_hook(hook)))
        self.apply(_register)
        try:
            self(*inputs)
            lines = ["%-30s %-30s %12s" % ("Layer (type)", "Output Shape", "Param #"), "=" * 74]
            total = 0
            for k, v in summary.items():
                lines.append("%-30s %-30s %12d" % (k, str(v["output_shape"][0] if v["output_shape"] else ""), v["n_params"]))
                total += v["n_params"]
            lines.append("=" * 74)
            lines.append("Total params (incl. shared/nested): %d" % total)
            print("\n".join(lines))
        finally:
            for _blk, handle in hooks:
                handle.detach()


class HookHandle(object):
    """Removable registration of a forward (pre-)hook (reference: python/mxnet/gluon/utils.py HookHandle)."""
    _next_id = [0]

    def __init__(self):
        self._hooks_dict_ref = None
        self._id = None

    def attach(self, hooks_dict, hook):
        import weakref
        assert not self._hooks_dict_ref, "The same handle cannot be attached twice."
        HookHandle._next_id[0] += 1
        self._id = HookHandle._next_id[0]
        hooks_dict[self._id] = hook
        self._hooks_dict_ref = weakref.ref(hooks_dict)

    def detach(self):
        hooks_dict = self._hooks_dict_ref() if self._hooks_dict_ref else None
        if hooks_dict is not None and self._id in hooks_dict:
            del hooks_dict[self._id]

    def __enter__(self):
        return self

    def __exit__(self, ptype, value, trace):
        self.detach()


class HybridBlock(Block):
    def __init__(self, prefix=None, params=None):
        super(HybridBlock, self).__init__(prefix=prefix, params=params)
        self._active = False
        self._flags = []
        self._cached_graph = ()

    def __setattr__(self, name, value):
        super(HybridBlock, self).__setattr__(name, value)
        if isinstance(value, HybridBlock):
            self._clear_cached_op()

    def _clear_cached_op(self):
        self._cached_graph = ()

    def register_child(self, block, name=None):
        if not isinstance(block, HybridBlock):
            raise ValueError("Children of HybridBlock must also be HybridBlock, but %s has type %s. If you are using "
                             "Sequential, please try HybridSequential instead." % (str(block), str(type(block))))
        super(HybridBlock, self).register_child(block, name)
        self._clear_cached_op()

    def hybridize(self, active=True, **kwargs):
        self._active = active
        self._flags = list(kwargs.items())
        self._clear_cached_op()
        super(HybridBlock, self).hybridize(active, **kwargs)

    def cast(self, dtype):
        self._clear_cached_op()
        super(HybridBlock, self).cast(dtype)

    def infer_shape(self, *args):
        """Layers with deferred parameters override ``_infer_param_shapes(*args)``."""
        self._infer_param_shapes(*args)

    def _infer_param_shapes(self, *args):
        pass

    def _deferred_infer(self, *args):
        self._infer_param_shapes(*args)
        for p in self._reg_params.values():
            p._finish_deferred_init()

    def forward(self, x, *args):
        if isinstance(x, NDArray):
            ctx = x.context
            try:
                params = {i: j._raw_data(ctx) for i, j in self._reg_params.items()}
            except DeferredInitializationError:
                self._deferred_infer(x, *args)
                params = {i: j._raw_data(ctx) for i, j in self._reg_params.items()}
            return self.hybrid_forward(nd, x, *args, **params)
        from .. import symbol
        assert isinstance(x, symbol.Symbol), "HybridBlock requires the first argument to forward be either Symbol " \
                                             "or NDArray, but got %s" % type(x)
        params = {i: j.var() for i, j in self._reg_params.items()}
        with self.name_scope():
            return self.hybrid_forward(symbol, x, *args, **params)

    def hybrid_forward(self, F, x, *args, **kwargs):
        raise NotImplementedError

    def _get_graph(self, *args):
        if not self._cached_graph:
            from .. import symbol
            args, self._in_format = _flatten(args, "input")
            if len(args) > 1:
                inputs = [symbol.var("data%d" % i) for i in range(len(args))]
            else:
                inputs = [symbol.var("data")]
            grouped_inputs = _regroup(inputs, self._in_format)[0]
            out = self._trace(symbol, grouped_inputs)
            out, self._out_format = _flatten(out, "output")
            self._cached_graph = inputs, symbol.Group(out) if len(out) > 1 else out[0]
        return self._cached_graph

    def _trace(self, symbol, grouped_inputs):
        if isinstance(grouped_inputs, (list, tuple)):
            return self.forward(*grouped_inputs)
        return self.forward(grouped_inputs)

    def export(self, path, epoch=0):
        """Write ``path-symbol.json`` and ``path-%04d.params`` with arg:/aux: prefixes (block.py:880-901)."""
        if not self._cached_graph:
            raise RuntimeError("Please first call block.hybridize() and then run forward with this block at least "
                               "once before calling export.")
        sym = self._cached_graph[1]
        sym.save("%s-symbol.json" % path)
        arg_names = set(sym.list_arguments())
        aux_names = set(sym.list_auxiliary_states())
        arg_dict = {}
        for name, param in self.collect_params().items():
            if name in arg_names:
                arg_dict["arg:%s" % name] = param._reduce()
            else:
                assert name in aux_names, "Parameter %s is not in the traced graph" % name
                arg_dict["aux:%s" % name] = param._reduce()
        _ser.save("%s-%04d.params" % (path, epoch), arg_dict)

    def __call__(self, *args):
        out = super(HybridBlock, self).__call__(*args)
        if self._active and not self._cached_graph and args and isinstance(args[0], NDArray):
            try:
                self._get_graph(*args)
            except Exception:  # noqa: BLE001  -- tracing is only needed for export(); never break training
                self._cached_graph = ()
        return out


class SymbolBlock(HybridBlock):
    """Construct a block from a Symbol graph (block.py:903-1040), e.g. a model loaded with ``imports``."""

    def _clear_cached_op(self):
        # the "cached graph" of a SymbolBlock IS its definition: hybridize() / re-parenting must not drop it
        self._cached_graph = getattr(self, "_defining_graph", ())

    @staticmethod
    def imports(symbol_file, input_names, param_file=None, ctx=None):
        from .. import symbol
        sym = symbol.load(symbol_file)
        if isinstance(input_names, str):
            input_names = [input_names]
        inputs = [symbol.var(i) for i in input_names]
        ret = SymbolBlock(sym, inputs)
        if param_file is not None:
            ret.collect_params().load(param_file, ctx=ctx)
        return ret

    def __init__(self, outputs, inputs, params=None):
        super(SymbolBlock, self).__init__(prefix=None, params=None)
        self._prefix = ""
        self._params = ParameterDict("", params)
        from .. import symbol
        if isinstance(inputs, symbol.Symbol) and len(inputs.list_outputs()) == 1:
            inputs = [inputs]
        if isinstance(outputs, (list, tuple)) and len(outputs) == 1:
            outputs = outputs[0]
        syms, self._in_format = _flatten(inputs, "input")
        out, self._out_format = _flatten(outputs, "output")
        out = symbol.Group(out) if len(out) > 1 else out[0]
        input_names = set(i.name for i in syms)
        # a SymbolBlock cannot own sparse parameters (block.py of the reference asserts on the inferred storage types)
        for node in out._topo():
            if node.op is None and node.name not in input_names:
                st = node.var_attrs.get("__storage_type__", "default")
                st = {"0": "default", "1": "row_sparse", "2": "csr"}.get(str(st), st)
                assert st == "default", \
                    "SymbolBlock doesn't support Parameter '%s' because its storage type is '%s'." % (node.name, st)
        for i in out.list_arguments():
            if i not in input_names:
                self.params.get(i, allow_deferred_init=True)
        for i in out.list_auxiliary_states():
            if i not in input_names:
                self.params.get(i, grad_req="null", allow_deferred_init=True)
        self._cached_graph = syms, out
        self._defining_graph = (syms, out)
        self._input_names = [i.name for i in syms]
        self._reg_params = {}

    def forward(self, x, *args):
        from .. import symbol
        if isinstance(x, NDArray):
            inputs = [x] + list(args)
            feed = dict(zip(self._input_names, inputs))
            ctx = x.context
            for name, p in self.params.items():
                try:
                    feed[name] = p._raw_data(ctx)
                except DeferredInitializationError:
                    sym_ = self._cached_graph[1]
                    arg_s, _o, aux_s = sym_.infer_shape_partial(**{n: a.shape for n, a in zip(self._input_names, inputs)})
                    shapes = dict(zip(sym_.list_arguments(), arg_s))
                    shapes.update(zip(sym_.list_auxiliary_states(), aux_s))
                    for pname, shp in shapes.items():
                        if pname in self.params._params and shp:
                            self.params[pname].shape = shp
                            self.params[pname]._finish_deferred_init()
                    feed[name] = p._raw_data(ctx)
            outs = self._cached_graph[1].eval_nd(feed)
            return outs[0] if len(outs) == 1 else outs
        return self._cached_graph[1]

    def hybrid_forward(self, F, x, *args, **kwargs):
        raise NotImplementedError
