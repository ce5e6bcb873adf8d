"""Executor: a Symbol bound to argument / gradient / aux arrays on one context
(reference: python/mxnet/executor.py, src/executor/graph_executor.cc).

forward() evaluates the graph with the shared op registry under the autograd scope; backward() runs torch
autograd from the outputs (loss ops such as SoftmaxOutput ignore head gradients exactly like the
reference).  Memory planning / bulk segments of GraphExecutor are unnecessary here: torch's caching
allocator and stream ordering play that role.
"""
import numpy as np
import torch

from .base import MXNetError
from .context import Context, cpu
from . import ndarray as nd
from .ndarray import NDArray
from . import autograd

__all__ = ["Executor"]

_NON_DIFFERENTIABLE = {"batch_take"}
_FWD_STAMP = [0]


class Executor(object):
    def __init__(self, symbol, ctx, args, args_grad=None, grad_req="write", aux_states=None, group2ctx=None):
        self._symbol = symbol
        self._ctx = ctx if isinstance(ctx, Context) else cpu()
        self._group2ctx = group2ctx
        arg_names = symbol.list_arguments()
        aux_names = symbol.list_auxiliary_states()
        if isinstance(args, dict):
            self.arg_dict = {n: args[n] for n in arg_names}
        else:
            assert len(args) == len(arg_names), "Length of args does not match the number of arguments"
            self.arg_dict = dict(zip(arg_names, args))
        if aux_states is None:
            aux_states = {}
        if isinstance(aux_states, dict):
            self.aux_dict = {n: aux_states[n] for n in aux_names}
        else:
            self.aux_dict = dict(zip(aux_names, aux_states))
        if isinstance(grad_req, str):
            self._grad_req = {n: grad_req for n in arg_names}
        elif isinstance(grad_req, (list, tuple)):
            self._grad_req = dict(zip(arg_names, grad_req))
        else:
            self._grad_req = {n: grad_req.get(n, "null") for n in arg_names}
        if args_grad is None:
            args_grad = {}
        # several DISTINCT variables may carry the same name (positional binding keeps them apart in the reference): such
        # graphs are fed by node identity, and the by-name dictionaries show the last array of each name
        self._by_node = None
        dup = len(set(arg_names)) != len(arg_names) and not isinstance(args, dict)
        if dup:
            var_nodes = [nd_ for nd_ in symbol._topo() if nd_.op is None and not nd_.is_aux]
            self._by_node = {id(nd_): a for nd_, a in zip(var_nodes, args)}
        pos_args = list(args) if dup else [self.arg_dict[n] for n in arg_names]
        if isinstance(args_grad, dict):
            pos_grads = [args_grad.get(n) for n in arg_names]
        else:
            pos_grads = list(args_grad) + [None] * (len(arg_names) - len(args_grad))
        pos_req = list(grad_req) if (dup and isinstance(grad_req, (list, tuple))) else [self._grad_req.get(n, "null") for n in arg_names]
        self.grad_dict = {}
        self._grad_cast = {}
        self._grad_pairs = []               # (argument array, bound gradient array, request) per position
        self.grad_arrays = []
        self._sparse_mirrors = {}
        sparse_grad_names = set()
        for node in symbol._topo():
            if (node.op == "Embedding" and str(node.attrs.get("sparse_grad", False)) in ("True", "1", "true")) \
                    or node.op == "_contrib_SparseEmbedding":
                if len(node.inputs) > 1 and node.inputs[1][0].op is None:
                    sparse_grad_names.add(node.inputs[1][0].name)
        for n, a, g, req in zip(arg_names, pos_args, pos_grads, pos_req):
            if req == "null":
                self.grad_arrays.append(None)
                continue
            if not a._t.is_floating_point():
                self._grad_req[n] = "null"
                self.grad_arrays.append(None)
                continue
            a.attach_grad(req)
            if g is not None and (g._t.dtype != a._t.dtype or g._t.device != a._t.device):
                # a gradient array of another dtype: autograd accumulates in the argument's dtype, backward() casts into it
                self._grad_cast[id(a)] = g
            elif g is not None:
                a._t.grad = g._t
            else:
                g = NDArray(a._t.grad, a.context)
            self._grad_pairs.append((a, g, req))
            if n in sparse_grad_names and n not in (args_grad if isinstance(args_grad, dict) else {}) and not dup:
                # Embedding(sparse_grad=True) / SparseEmbedding weights: the executor-owned gradient is handed out as a
                # row_sparse array (rows = the looked-up ids), refreshed from the dense accumulator after every backward()
                from .ndarray import sparse as _sp
                mirror = _sp.zeros("row_sparse", a.shape, ctx=a.context, dtype=a.dtype)
                self._sparse_mirrors[n] = (a, mirror)
                g = mirror
            self.grad_dict[n] = g
            self.grad_arrays.append(g)
        self.arg_arrays = pos_args
        if self._grad_pairs:
            # operators without a gradient definition cannot sit on a path a requested gradient takes (nnvm gradient pass:
            # "Operator X is non-differentiable because it didn't register FGradient attribute")
            _reach, ops = self._grad_reach()
            for node in ops:
                if node.op in _NON_DIFFERENTIABLE:
                    raise MXNetError("Operator %s is non-differentiable because it didn't register FGradient attribute." % node.op)
        self.aux_arrays = [self.aux_dict[n] for n in aux_names]
        self._outputs = []
        self._output_names = symbol.list_outputs()
        self._monitor_callback = None

    @property
    def outputs(self):
        """Output arrays.  Before the first forward() they are zero arrays of the inferred shapes / dtypes (the reference
        allocates them at bind time, so scripts read ``exe.outputs[0].shape`` right after ``bind``)."""
        if not self._outputs:
            try:
                shapes = {n: a.shape for n, a in list(self.arg_dict.items()) + list(self.aux_dict.items())}
                from .base import dtype_name
                dtypes = {n: dtype_name(a._t.dtype) for n, a in list(self.arg_dict.items()) + list(self.aux_dict.items())}
                _v, out_shapes, _d, out_types = self._symbol._infer(shapes, dtypes)
                if all(s is not None for s in out_shapes):
                    self._outputs = [nd.zeros(s, ctx=self._ctx, dtype=t) for s, t in zip(out_shapes, out_types)]
            except Exception:  # noqa: BLE001 -- best effort; forward() fills the real values
                pass
        for ref in self.__dict__.get("_children", ()):
            c = ref()
            if c is not None and c._outputs and getattr(c, "_fwd_stamp", 0) > getattr(self, "_fwd_stamp", 0):
                for mine, theirs in zip(self._outputs, c._outputs):
                    if mine.shape[1:] == theirs.shape[1:] and theirs.shape[0] <= mine.shape[0] and mine._t.dtype == theirs._t.dtype:
                        with torch.no_grad():
                            mine._t[:theirs.shape[0]].copy_(theirs._t.detach())
                self._fwd_stamp = c._fwd_stamp
        return self._outputs

    @outputs.setter
    def outputs(self, value):
        self._outputs = value

    @property
    def output_dict(self):
        return dict(zip(self._output_names, self.outputs))

    @staticmethod
    def simple_bind(symbol, ctx, grad_req="write", type_dict=None, group2ctx=None, shapes=None):
        known = set(symbol.list_inputs())
        shapes = {k: v for k, v in (shapes or {}).items() if k in known}      # stray keyword arguments are ignored
        type_dict = type_dict or {}
        vshape, _out, vdtype, _ = symbol._infer(shapes, type_dict)
        args, aux = {}, {}
        # model parallelism by placement: a variable created under AttrScope(ctx_group=...) lives on group2ctx[group]
        # (graph_executor.cc AssignContext); everything else on the default context
        place = {}
        if group2ctx:
            for node in symbol._topo():
                if node.op is None:
                    g = node.var_attrs.get("__ctx_group__")
                    if g in group2ctx:
                        place[node.name] = group2ctx[g] if not isinstance(group2ctx[g], (list, tuple)) else group2ctx[g][0]
        for n in symbol.list_arguments():
            if n not in vshape:
                raise MXNetError("simple_bind: cannot infer shape of argument %s" % n)
            args[n] = nd.zeros(vshape[n], ctx=place.get(n, ctx), dtype=type_dict.get(n, vdtype.get(n, "float32")))
        for n in symbol.list_auxiliary_states():
            aux[n] = nd.zeros(vshape[n], ctx=place.get(n, ctx), dtype="float32")
        if isinstance(grad_req, str):
            req = {n: ("null" if n in shapes and grad_req != "null" and False else grad_req) for n in args}
        else:
            req = grad_req
        return Executor(symbol, ctx, args, None, req, aux, group2ctx)

    def forward(self, is_train=False, **kwargs):
        for name, array in kwargs.items():
            if name not in self.arg_dict:
                raise TypeError("Unknown argument %s" % name)
            if not isinstance(array, NDArray):
                array = nd.array(array, ctx=self._ctx)
            if self.arg_dict[name].shape == array.shape:
                self.arg_dict[name][:] = array
            else:
                raise ValueError("Shape not match! Argument %s, need: %s, received: %s" % (
                    name, str(self.arg_dict[name].shape), str(array.shape)))
        feed = dict(self.arg_dict)
        feed.update(self.aux_dict)
        if self._by_node is not None:
            feed["__by_node__"] = self._by_node
        mon = self._make_monitor() if self._monitor_callback is not None else None
        if is_train or self.grad_dict:
            # an executor bound with gradient arrays can always run backward(), also after forward(is_train=False): the
            # mode only selects training / inference behaviour of the operators (graph_executor.cc RunOps)
            with autograd.record(train_mode=bool(is_train)):
                self.outputs = self._symbol.eval_nd(feed, self._group2ctx, monitor=mon)
        else:
            with autograd.pause(train_mode=False):
                self.outputs = self._symbol.eval_nd(feed, self._group2ctx, monitor=mon)
        self._ran_forward = True
        _FWD_STAMP[0] += 1
        self._fwd_stamp = _FWD_STAMP[0]
        self._last_is_train = bool(is_train)
        return self.outputs

    def _make_monitor(self):
        """Adapter between Symbol.eval_nd's (name, array, when) and the user's monitor callback: operator outputs are
        reported at once; for the backward pass a hook on every operator input reports the gradient that reaches it
        (inputs that carry no gradient -- data, labels -- are reported as zeros after backward, like the reference's backward
        graph which materialises them)."""
        cb = self._monitor_callback
        self._mon_nograd = []

        def mon(name, arr, when):
            if when == "forward":
                cb(name, arr)
                return
            t = arr._t if hasattr(arr, "_t") else None
            if t is None:
                return
            if t.requires_grad and t.is_floating_point():
                ctx = arr.context
                t.register_hook(lambda g, name=name, ctx=ctx: (cb(name, NDArray(g.detach(), ctx)), None)[1])
            else:
                self._mon_nograd.append((name, arr))
        return mon

    def backward(self, out_grads=None, is_train=True):
        if out_grads is None:
            out_grads = []
        elif isinstance(out_grads, NDArray):
            out_grads = [out_grads]
        elif isinstance(out_grads, dict):
            out_grads = [out_grads[k] for k in self._output_names]
        if not getattr(self, "_ran_forward", False) and self.grad_dict:
            self.forward(is_train=is_train)      # backward() straight after bind: the graph has to be evaluated once
        heads, hgrads = [], []
        for i, o in enumerate(self.outputs):
            if not o._t.requires_grad:
                continue
            heads.append(o)
            hgrads.append(out_grads[i] if i < len(out_grads) else None)
        epoch = autograd._bw_epoch[0]
        if heads and any(getattr(h._t, "_mx_graph_freed", False) for h in heads):
            # backward() twice on one forward(): the reference's static graph allows it; the tape here was consumed by the
            # first call, so the forward pass is replayed on the same inputs (stateful operators -- BatchNorm statistics,
            # dropout masks -- see a second forward)
            self.forward(is_train=getattr(self, "_last_is_train", True))
            heads = [o for o in self.outputs if o._t.requires_grad]
        if heads:
            autograd.backward(heads, hgrads)
        # 'write' means every bound gradient array is written by every backward(): arguments no gradient reached (inputs of
        # integer-valued operators, disconnected branches) read zero, not the previous contents.  "Not reached" = no path from
        # a differentiable head in the graph, or the array's write hook did not fire in this pass; arrays whose hook was
        # taken over by a fused update path (Module._maybe_fuse clears its gradient arena itself) are left alone.
        head_ids = tuple(i for i, o in enumerate(self.outputs) if o._t.requires_grad)
        cache = self.__dict__.setdefault("_reach_cache", {})
        if head_ids not in cache:
            cache[head_ids] = self._grad_reach(head_ids)[0]
        reach = cache[head_ids]
        for (a, g, req), n in zip(self._grad_pairs, self._grad_pair_names()):
            leaf = getattr(a, "_dense_leaf", None)                # sparse arguments accumulate in their dense leaf
            leaf = a if leaf is None else leaf
            unreachable = (not heads) or (n is not None and n not in reach)
            missed = getattr(leaf, "_grad_hook", None) is not None and getattr(leaf, "_bw_epoch", -1) <= epoch
            if req == "write" and (unreachable or missed) and a._t.grad is not None:
                a._t.grad.zero_()
            live = a._t.grad
            if live is None:
                continue
            if id(a) in self._grad_cast:
                with torch.no_grad():
                    g._t.copy_(live)
            elif live is not g._t:
                g._t = live                       # keep the handles pointing at the live tensors
        for n, (a, mirror) in self._sparse_mirrors.items():
            if a._t.grad is not None:
                from .ndarray import sparse as _sp
                _sp.cast_storage(NDArray(a._t.grad.detach(), a.context), "row_sparse").copyto(mirror)
        # a failed operator's exception travels with everything computed from it: the gradients too
        for o in self.outputs:
            exc = getattr(o._t, "_mx_exc", None)
            if exc is not None:
                for _a, g, _r in self._grad_pairs:
                    g._t._mx_exc = exc
                break
        if self._monitor_callback is not None:
            for name, arr in getattr(self, "_mon_nograd", []):
                self._monitor_callback(name, nd.zeros(arr.shape, ctx=arr.context, dtype=arr.dtype))
            self._mon_nograd = []

    def set_monitor_callback(self, callback):
        self._monitor_callback = callback

    def copy_params_from(self, arg_params, aux_params=None, allow_extra_params=False):
        for name, array in arg_params.items():
            if name in self.arg_dict:
                dst = self.arg_dict[name]
                with torch.no_grad():
                    dst._t.copy_(array._t.to(dst._t.dtype))
            elif not allow_extra_params:
                raise ValueError("Find name \"%s\" that is not in the arguments" % name)
        if aux_params is None:
            return
        for name, array in aux_params.items():
            if name in self.aux_dict:
                dst = self.aux_dict[name]
                with torch.no_grad():
                    dst._t.copy_(array._t.to(dst._t.dtype))
            elif not allow_extra_params:
                raise ValueError("Find name %s that is not in the auxiliary states" % name)

    def reshape(self, partial_shaping=False, allow_up_sizing=False, **kwargs):
        vshape, _o, vdtype, _ = self._symbol._infer(kwargs)
        new_args, new_aux = {}, {}
        for n, a in self.arg_dict.items():
            shp = tuple(vshape.get(n, a.shape))
            if shp == a.shape:
                new_args[n] = a
                continue
            size = int(np.prod(shp)) if shp else 1
            if size <= a.size and a._t.is_contiguous():
                # smaller view of the SAME memory, like the reference (graph_executor.cc Reshape shares the arrays)
                new_args[n] = NDArray(a._t.detach().reshape(-1)[:size].view(shp), a.context)
            elif size > a.size and not allow_up_sizing:
                raise AssertionError("New shape of arg: %s is larger than original. First making a big executor and then down "
                                     "sizing it is more efficient than the reverse. If you really want to up size, set "
                                     "allow_up_sizing=True to enable allocation of new arrays." % n)
            else:
                new_args[n] = nd.zeros(shp, ctx=self._ctx, dtype=a.dtype)
        for n, a in self.aux_dict.items():
            new_aux[n] = a
        child = Executor(self._symbol, self._ctx, new_args, None, self._grad_req, new_aux, self._group2ctx)
        # the reference's reshaped executor works in the parent's memory: outputs of the child are visible in the leading
        # part of the parent's output arrays.  Mirrored lazily, when the parent's outputs are read after the child ran.
        import weakref
        self.__dict__.setdefault("_children", []).append(weakref.ref(child))
        return child

    def _rebind_arrays(self):
        """After arg_dict / grad_dict entries were replaced from outside (parameter sharing between bucket executors):
        refresh the per-position bookkeeping backward() works from."""
        if self._by_node is not None:
            return
        names = self._symbol.list_arguments()
        self.arg_arrays = [self.arg_dict[n] for n in names]
        self.grad_arrays = [self.grad_dict.get(n) for n in names]
        keep = {id(a): (a, g, r) for a, g, r in self._grad_pairs}
        self._grad_pairs = [keep[id(self.arg_dict[n])] if n in self._sparse_mirrors and id(self.arg_dict[n]) in keep
                            else (self.arg_dict[n], self.grad_dict[n], self._grad_req.get(n, "null"))
                            for n in names if n in self.grad_dict and self._grad_req.get(n, "null") != "null"]
        self.__dict__.pop("_pair_names", None)
        self._grad_cast = {id(a): g for a, g, _r in self._grad_pairs
                           if g._t.dtype != a._t.dtype or g._t.device != a._t.device}

    def _grad_pair_names(self):
        """Argument name of every entry of ``_grad_pairs`` (None when names are ambiguous: duplicate-named variables)."""
        names = self.__dict__.get("_pair_names")
        if names is None or len(names) != len(self._grad_pairs):
            if self._by_node is not None:
                names = [None] * len(self._grad_pairs)
            else:
                by_id = {id(a): n for n, a in self.arg_dict.items()}
                names = [by_id.get(id(a)) for a, _g, _r in self._grad_pairs]
            self._pair_names = names
        return names

    def _grad_reach(self, head_ids=None):
        """Names of the arguments a gradient can reach from the outputs (zero-gradient propagation stops at BlockGrad /
        stop_gradient like the reference's gradient pass) and the operator nodes on those paths."""
        reach, ops, seen = set(), [], set()
        outs = self._symbol._outputs
        stack = [n for k, (n, _i) in enumerate(outs) if head_ids is None or k in head_ids]
        while stack:
            n = stack.pop()
            if id(n) in seen:
                continue
            seen.add(id(n))
            if n.op is None:
                reach.add(n.name)
                continue
            if n.op in ("BlockGrad", "stop_gradient"):
                continue
            ops.append(n)
            stack.extend(i for i, _ in n.inputs)
        return reach, ops

    def debug_str(self):
        """Graph listing plus the memory the bound graph needs (graph_executor.cc debug string: "Total N MB allocated").
        Inference-only graphs (no gradient requested, or none can reach an argument) recycle intermediates; a training
        graph keeps every intermediate for the backward pass and owns the gradient arrays."""
        lines = ["Symbol Outputs:"] + ["\toutput[%d]=%s" % (i, n) for i, n in enumerate(self._output_names)]
        size = lambda a: int(np.prod(a.shape)) * a._t.element_size() if hasattr(a, "_t") else 0   # noqa: E731
        total = sum(size(a) for a in self.arg_dict.values()) + sum(size(a) for a in self.aux_dict.values())
        reach, _ops = self._grad_reach()
        training = any(self._grad_req.get(n, "null") != "null" for n in reach if n in self.arg_dict)
        try:
            internals = self._symbol.get_internals()
            shapes = {n: a.shape for n, a in list(self.arg_dict.items()) + list(self.aux_dict.items())}
            _a, outs, _x = internals.infer_shape(**shapes)
            names = internals.list_outputs()
            inter = [int(np.prod(s)) * 4 for n, s in zip(names, outs) if s is not None and n not in self.arg_dict and n not in self.aux_dict]
        except Exception:  # noqa: BLE001 -- best effort report
            inter = []
        for node in self._symbol._topo():
            if node.op is not None:
                lines.append("Op:%s, Name=%s" % (node.op, node.name))
        if training:
            total += sum(inter) + sum(size(self.arg_dict[n]) for n in reach if n in self.arg_dict and self._grad_req.get(n, "null") != "null")
        else:
            total += sum(sorted(inter)[-2:])               # ping-pong buffers
        lines.append("Total %d MB allocated" % (total >> 20))
        return "\n".join(lines)
