"""Symbolic graph API (reference: python/mxnet/symbol/symbol.py, 3rdparty/tvm/nnvm graph IR + passes).

A ``Symbol`` is a list of output entries of a small DAG of op nodes.  Ops are the same functions as in
``mx.nd`` (one registry), so the graph is *executed* by running those functions in topological order on
NDArrays (``eval_nd`` / ``Executor``) with torch autograd providing gradients -- replacing the reference's
NNVM gradient / PlanMemory / bulk-exec passes (src/executor/graph_executor.cc).  Shape/type inference runs
the graph on ``meta`` tensors after filling in parameter shapes with per-op rules (FullyConnected,
Convolution, BatchNorm, ...), the counterpart of infer_shape_type.cc.  ``tojson`` writes the NNVM JSON layout
(nodes / arg_nodes / node_row_ptr / heads) so ``-symbol.json`` checkpoints round-trip.
"""
import copy
import json
import threading

import numpy as np
import torch

from ..base import MXNetError, torch_dtype, np_dtype, dtype_name
from ..context import cpu, Context
from .. import ndarray as nd
from ..ndarray import NDArray
from ..ndarray import op as _ndop
from .. import autograd
from ..base import getenv

__all__ = ["Symbol", "var", "Variable", "Group", "load", "load_json", "zeros", "ones"]

# ops that own parameters: op -> (list of (arg_name, is_aux)), created as variables when not passed explicitly
_PARAM_ARGS = {
    "FullyConnected": [("weight", False), ("bias", False)],
    "Convolution": [("weight", False), ("bias", False)],
    "Deconvolution": [("weight", False), ("bias", False)],
    "BatchNorm": [("gamma", False), ("beta", False), ("moving_mean", True), ("moving_var", True)],
    "LayerNorm": [("gamma", False), ("beta", False)],
    "InstanceNorm": [("gamma", False), ("beta", False)],
    "Embedding": [("weight", False)],
    "_contrib_SparseEmbedding": [("weight", False)],
    "LeakyReLU": [("gamma", False)],
    "BatchNorm_v1": [("gamma", False), ("beta", False), ("moving_mean", True), ("moving_var", True)],
    "CuDNNBatchNorm": [("gamma", False), ("beta", False), ("moving_mean", True), ("moving_var", True)],
    "_contrib_SyncBatchNorm": [("gamma", False), ("beta", False), ("moving_mean", True), ("moving_var", True)],
    "Convolution_v1": [("weight", False), ("bias", False)],
    "_contrib_DeformableConvolution": [("offset", False), ("weight", False), ("bias", False)],
    "IdentityAttachKLSparseReg": [("moving_avg", True)],
}


def _fused_resnet_args(attrs):
    from ..models.resnet import ResNetSpec, resnet_param_spec
    spec = ResNetSpec(int(attrs.get("num_layers", 50)), classes=int(attrs.get("classes", 1000)), version=str(attrs.get("version", "v1b")),
                      in_channels=int(attrs.get("in_channels", 3)), thumbnail=_truthy(attrs.get("thumbnail", False)))
    return spec, resnet_param_spec(spec)


# ops whose parameter list depends on their attributes: op -> fn(attrs) -> [(arg_name, is_aux)]
_DYNAMIC_PARAM_ARGS = {
    "_contrib_FusedResNet": lambda attrs: [(n, aux) for n, _s, _k, aux in _fused_resnet_args(attrs)[1]],
}


def _param_args(op, attrs):
    if op in _DYNAMIC_PARAM_ARGS:
        return _DYNAMIC_PARAM_ARGS[op](attrs)
    return _PARAM_ARGS.get(op, [])


def _has_param_args(op):
    return op in _PARAM_ARGS or op in _DYNAMIC_PARAM_ARGS


_LABEL_OPS = {"SoftmaxOutput": "label", "Softmax": "label", "LinearRegressionOutput": "label",
              "LogisticRegressionOutput": "label", "MAERegressionOutput": "label", "SVMOutput": "label"}
_MULTI_OUT = {"split": "num_outputs", "SliceChannel": "num_outputs"}
_FIXED_MULTI_OUT = {"_contrib_MultiBoxTarget": 3, "linalg_gelqf": 2, "_linalg_gelqf": 2, "linalg_syevd": 2, "_linalg_syevd": 2,
                    "_contrib_quantize": 3, "_contrib_requantize": 3, "_contrib_bipartite_matching": 2, "histogram": 2,
                    "_histogram": 2, "_contrib_quantized_fully_connected": 3, "_contrib_quantized_conv": 3,
                    "_contrib_quantized_pooling": 3, "_contrib_quantized_flatten": 3}


def _truthy(v):
    return v in (True, 1, "1", "True", "true")


def _custom_prop(attrs):
    """CustomOpProp of a Custom node (operator.py registry); its constructor takes the node's parameters as strings."""
    from .. import operator as _custom
    return _custom.get_prop(attrs.get("op_type"), {k: v for k, v in attrs.items() if k != "op_type" and not k.startswith("__")})


def _num_outputs(op, attrs):
    """Number of outputs of an operator node (nnvm FNumOutputs of the reference registrations)."""
    if op in _MULTI_OUT:
        return int(attrs.get(_MULTI_OUT[op], 1))
    if op in _FIXED_MULTI_OUT:
        return _FIXED_MULTI_OUT[op]
    if op == "Custom":
        return len(_custom_prop(attrs).list_outputs())
    if op == "RNN":
        if not _truthy(attrs.get("state_outputs", False)):
            return 1
        return 3 if str(attrs.get("mode", "lstm")) == "lstm" else 2
    if op == "topk":
        return 2 if str(attrs.get("ret_typ", "indices")) == "both" else 1
    if op in ("BatchNorm", "BatchNorm_v1", "CuDNNBatchNorm", "LayerNorm"):
        return 3 if _truthy(attrs.get("output_mean_var", False)) else 1
    if op == "_sample_multinomial":
        return 2 if _truthy(attrs.get("get_prob", False)) else 1
    if op in ("_contrib_Proposal", "_contrib_MultiProposal"):
        return 2 if _truthy(attrs.get("output_score", False)) else 1
    return 1


# attribute keys the reference stores with double underscores whatever the user wrote (c_api_symbolic.cc kHiddenKeys)
_HIDDEN_ATTR_KEYS = ("ctx_group", "lr_mult", "wd_mult", "force_mirroring", "mirror_stage")


def _attr_key(k):
    return "__%s__" % k if k in _HIDDEN_ATTR_KEYS else k


def _put_attr(d, k, v, override=True):
    """Store a user attribute; hidden keys are visible under both spellings (key and __key__)."""
    keys = (k, "__%s__" % k) if k in _HIDDEN_ATTR_KEYS else (k,)
    for kk in keys:
        if override or kk not in d:
            d[kk] = v


class _NameManager(object):
    """Bridge to mx.name.NameManager (per-thread, scoped by ``with``) plus the Gluon block prefix."""

    @classmethod
    def get(cls, hint):
        from ..name import NameManager
        return NameManager._get_current().get(None, hint)

    @classmethod
    def resolve(cls, name, hint):
        # inside a Gluon name scope the current NameManager is a ``Prefix(block.prefix)`` pushed by _BlockScope.__enter__
        # (python/mxnet/gluon/block.py:_BlockScope of the reference), so the block prefix is already part of the name
        from ..name import NameManager
        return NameManager._get_current().get(name, hint)

    @classmethod
    def prefix(cls):
        from ..gluon.block import _BlockScope
        cur = getattr(_BlockScope._current, "value", None)
        return cur._block.prefix if cur is not None else ""


class _Node(object):
    __slots__ = ("op", "name", "attrs", "inputs", "is_aux", "num_outputs", "var_attrs")

    def __init__(self, op, name, attrs=None, inputs=None, num_outputs=1):
        self.op = op            # None for variables
        self.name = name
        self.attrs = attrs or {}
        self.inputs = inputs or []      # list of (node, out_index)
        self.is_aux = False
        self.num_outputs = num_outputs
        self.var_attrs = {}


class Symbol(object):
    def __init__(self, outputs):
        self._outputs = list(outputs)   # list of (node, index)

    # ---- composition helpers -------------------------------------------------------------------
    @property
    def name(self):
        if len(self._outputs) != 1:
            return None
        return self._outputs[0][0].name

    def __repr__(self):
        if len(self._outputs) == 1:
            return "<Symbol %s>" % self.name
        return "<Symbol group [%s]>" % ", ".join(n.name for n, _ in self._outputs)

    def __iter__(self):
        return (self[i] for i in range(len(self._outputs)))

    def __len__(self):
        return len(self._outputs)

    def __getitem__(self, index):
        if isinstance(index, str):
            names = self.list_outputs()
            idx = [i for i, n in enumerate(names) if n == index]
            if len(idx) != 1:
                raise ValueError('There are multiple outputs with name "%s"' % index if idx else "Cannot find output %s" % index)
            index = idx[0]
        if isinstance(index, slice):
            return Symbol(self._outputs[index])
        return Symbol([self._outputs[index]])

    def _topo(self):
        order, seen = [], set()

        def visit(node):
            if id(node) in seen:
                return
            seen.add(id(node))
            for inp, _ in node.inputs:
                visit(inp)
            order.append(node)
        import sys
        sys.setrecursionlimit(max(10000, sys.getrecursionlimit()))
        for n, _ in self._outputs:
            visit(n)
        return order

    def list_arguments(self):
        return [n.name for n in self._topo() if n.op is None and not n.is_aux]

    def list_auxiliary_states(self):
        return [n.name for n in self._topo() if n.op is None and n.is_aux]

    def list_inputs(self):
        return [n.name for n in self._topo() if n.op is None]

    def list_outputs(self):
        out = []
        for n, i in self._outputs:
            if n.op is None:
                out.append(n.name)
            elif n.num_outputs > 1:
                out.append("%s_%s" % (n.name, _custom_prop(n.attrs).list_outputs()[i]) if n.op == "Custom" else "%s_output%d" % (n.name, i))
            else:
                out.append("%s_output" % n.name)
        return out

    def get_internals(self):
        outs = []
        for n in self._topo():
            for i in range(n.num_outputs):
                outs.append((n, i))
        return Symbol(outs)

    def get_children(self):
        """Inputs of the output node(s) as a grouped symbol; None for variables (symbol.py:get_children)."""
        inputs, seen = [], set()
        for n, _ in self._outputs:
            if id(n) not in seen:                  # several outputs of one node (SliceChannel) share their inputs
                seen.add(id(n))
                inputs.extend(n.inputs)
        return Symbol(inputs) if inputs else None

    def attr(self, key):
        n = self._outputs[0][0]
        v = n.var_attrs.get(key, n.attrs.get(key))
        if v is None and key in _HIDDEN_ATTR_KEYS:
            v = n.var_attrs.get("__%s__" % key)
        return _attr_str(v) if v is not None else None

    def list_attr(self):
        n = self._outputs[0][0]
        d = {k: str(v) for k, v in n.attrs.items()}
        d.update({k: str(v) for k, v in n.var_attrs.items()})
        return d

    def attr_dict(self):
        ret = {}
        for n in self._topo():
            d = {k: str(v) for k, v in n.attrs.items()}
            d.update({k: str(v) for k, v in n.var_attrs.items()})
            if d:
                ret[n.name] = d
        return ret

    def _set_attr(self, **kwargs):
        for k, v in kwargs.items():
            if not isinstance(v, str):
                raise ValueError("Set Attr only accepts string values")
        for k, v in kwargs.items():
            _put_attr(self._outputs[0][0].var_attrs, k, v)

    # ---- arithmetic ------------------------------------------------------------------------------
    def _bin(self, other, op, scalar_op, rev=False):
        if isinstance(other, Symbol):
            a, b = (other, self) if rev else (self, other)
            # the reference's symbol operators are the same-shape element-wise ops (_Plus, ...): shape inference may unify
            # the operands; evaluation here broadcasts, which is a superset
            return _make_op(op, [a, b], {"__same_shape__": True})
        return _make_op(scalar_op, [self], {"scalar": float(other), "reverse": rev})

    def __add__(self, o):
        return self._bin(o, "broadcast_add", "_plus_scalar")

    __radd__ = __add__

    def __sub__(self, o):
        return self._bin(o, "broadcast_sub", "_minus_scalar")

    def __rsub__(self, o):
        return self._bin(o, "broadcast_sub", "_minus_scalar", rev=True)

    def __mul__(self, o):
        return self._bin(o, "broadcast_mul", "_mul_scalar")

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._bin(o, "broadcast_div", "_div_scalar")

    def __rtruediv__(self, o):
        return self._bin(o, "broadcast_div", "_div_scalar", rev=True)

    __div__, __rdiv__ = __truediv__, __rtruediv__

    def __pow__(self, o):
        return self._bin(o, "broadcast_power", "_power_scalar")

    def __neg__(self):
        return self * -1.0

    def __gt__(self, o):
        return self._bin(o, "broadcast_greater", "_greater_scalar")

    def __lt__(self, o):
        return self._bin(o, "broadcast_lesser", "_lesser_scalar")

    def __ge__(self, o):
        return self._bin(o, "broadcast_greater_equal", "_greater_equal_scalar")

    def __le__(self, o):
        return self._bin(o, "broadcast_lesser_equal", "_lesser_equal_scalar")

    def __ne__(self, o):
        if isinstance(o, (Symbol, int, float)):
            return self._bin(o, "broadcast_not_equal", "_not_equal_scalar")
        return NotImplemented

    def __eq__(self, o):
        if isinstance(o, (Symbol, int, float)):
            return self._bin(o, "broadcast_equal", "_equal_scalar")
        return NotImplemented

    __hash__ = object.__hash__

    def __mod__(self, o):
        return self._bin(o, "broadcast_mod", "_mod_scalar")

    def __rmod__(self, o):
        return self._bin(o, "broadcast_mod", "_mod_scalar", rev=True) if isinstance(o, Symbol) else \
            _make_op("_rmod_scalar", [self], {"scalar": float(o)})

    def __rpow__(self, o):
        return _make_op("_rpower_scalar", [self], {"scalar": float(o)})

    # ---- things only an NDArray can do (python/mxnet/symbol/symbol.py raises NotImplementedForSymbol for these)
    def __bool__(self):
        from ..base import NotImplementedForSymbol
        raise NotImplementedForSymbol(self.__bool__, "bool")

    __nonzero__ = __bool__

    def _ndarray_only(fname, alias=None):   # noqa: N805
        def f(self, *args, **kwargs):
            from ..base import NotImplementedForSymbol
            raise NotImplementedForSymbol(getattr(Symbol, fname), alias, *args)
        f.__name__ = fname
        return f

    __iadd__ = _ndarray_only("__iadd__", "+=")
    __isub__ = _ndarray_only("__isub__", "-=")
    __imul__ = _ndarray_only("__imul__", "*=")
    __idiv__ = _ndarray_only("__idiv__", "/=")
    __itruediv__ = _ndarray_only("__itruediv__", "/=")
    __imod__ = _ndarray_only("__imod__", "%=")
    wait_to_read = _ndarray_only("wait_to_read")
    asnumpy = _ndarray_only("asnumpy")
    asscalar = _ndarray_only("asscalar")
    copy = _ndarray_only("copy")
    as_in_context = _ndarray_only("as_in_context")
    detach = _ndarray_only("detach")
    backward = _ndarray_only("backward")
    del _ndarray_only

    def astype(self, dtype):
        return _make_op("Cast", [self], {"dtype": dtype_name(dtype)})

    # ---- composition (symbol.py:375-470 of the reference) ------------------------------------------
    def __copy__(self):
        return load_json(self.tojson())

    def __call__(self, *args, **kwargs):
        """``net2(fc3_data=net1)``: returns a copy of this symbol whose free variables are replaced by other symbols."""
        s = self.__copy__()
        s._compose(*args, **kwargs)
        return s

    def _compose(self, *args, **kwargs):
        name = kwargs.pop("name", None)
        if args and kwargs:
            raise TypeError("compose only accept input Symbols either as positional or keyword arguments, not both")
        for a in list(args) + list(kwargs.values()):
            if not isinstance(a, Symbol):
                raise TypeError("Compose expect `Symbol` as arguments")
        free = [n for n in self._topo() if n.op is None and not n.is_aux]
        if args:
            if len(args) > len(free):
                raise MXNetError("compose: %d positional arguments for %d free variables" % (len(args), len(free)))
            repl = {id(n): a._outputs[0] for n, a in zip(free, args)}
        else:
            byname = {n.name: n for n in free}
            unknown = [k for k in kwargs if k not in byname]
            if unknown:
                raise MXNetError("compose: no free variable named %s (free variables: %s)" % (unknown, sorted(byname)))
            repl = {id(byname[k]): v._outputs[0] for k, v in kwargs.items()}
        for n in self._topo():
            n.inputs = [repl.get(id(inp), (inp, idx)) if idx == 0 else (inp, idx) for inp, idx in n.inputs]
        self._outputs = [repl.get(id(n), (n, i)) if i == 0 else (n, i) for n, i in self._outputs]
        if name and len(self._outputs) == 1:
            self._outputs[0][0].name = name
        self._fusion_cache = None

    def debug_str(self):
        lines = []
        for n in self._topo():
            if n.op is None:
                lines.append("Variable:%s" % n.name)
            else:
                lines.append("--------------------\nOp:%s, Name=%s\nInputs:" % (n.op, n.name))
                lines += ["\targ[%d]=%s(%d)" % (i, inp.name, idx) for i, (inp, idx) in enumerate(n.inputs)]
                if n.attrs:
                    lines.append("Attrs:")
                    lines += ["\t%s=%s" % (k, _attr_str(v)) for k, v in n.attrs.items() if not k.startswith("__kwarg")]
        return "Symbol Outputs:\n" + "\n".join("\toutput[%d]=%s(%d)" % (i, n.name, idx) for i, (n, idx) in enumerate(self._outputs)) \
            + "\n" + "\n".join(lines) + "\n"

    def gradient(self, wrt):
        raise MXNetError("Symbol.gradient is not supported (deprecated in the reference as well: use bind + backward, or autograd)")

    def get_backend_symbol(self, backend):
        """Graph partitioning for an accelerator backend (MKLDNN / TensorRT in the reference).  The fusion passes of this
        framework run inside the executor, so the symbol is returned unchanged."""
        return self

    def __getattr__(self, name):
        # method forms: sym.reshape(...), sym.sum(...), ...
        if name.startswith("_"):
            raise AttributeError(name)
        from . import _lookup_op
        fn = _lookup_op(name)
        if fn is None:
            raise AttributeError("Symbol has no attribute %r" % name)
        return lambda *a, **k: fn(self, *a, **k)

    def reshape(self, *shape, **kwargs):
        if len(shape) == 1 and isinstance(shape[0], (list, tuple)):
            shape = tuple(shape[0])
        if not shape:
            shape = tuple(kwargs["shape"])
        attrs = {"shape": tuple(shape)}
        if kwargs.get("reverse", False):
            attrs["reverse"] = True
        return _make_op("reshape", [self], attrs)

    # ---- inference -------------------------------------------------------------------------------
    def _infer(self, shapes, dtypes=None, partial=False):
        """Returns ({var_name: shape}, [out shapes], {var_name: dtype}, [out dtypes])."""
        shapes = {k: ((int(v),) if isinstance(v, (int, np.integer)) else tuple(v)) for k, v in shapes.items() if v is not None}
        dtypes = dict(dtypes or {})
        # partially known shapes (0 = unknown dimension) and variables without a shape: a bidirectional pre-pass completes what
        # the graph determines (shape_solver.py: NNVM's forward + backward InferShape sweeps); the forward evaluation below
        # then sees complete shapes
        order0 = self._topo()
        need = any(n.op is None and shapes.get(n.name) is None and
                   not ("__shape__" in n.var_attrs and all(d > 0 for d in n.var_attrs["__shape__"])) for n in order0) \
            or any(any(d == 0 for d in v) for v in shapes.values())
        if need:
            from .shape_solver import solve
            solved, solved_heads = solve(self, shapes, partial_heads=partial)
            shapes = {k: v for k, v in shapes.items() if all(d > 0 for d in v)}
            for k, v in solved.items():
                shapes.setdefault(k, v)
        else:
            solved_heads = None
        vals = {}
        vshape, vdtype = {}, {}
        order = self._topo()
        for node in order:
            if node.op is None:
                shp = shapes.get(node.name)
                if shp is None and "__shape__" in node.var_attrs:
                    s = tuple(node.var_attrs["__shape__"])
                    shp = s if all(d > 0 for d in s) else None
                dt = dtypes.get(node.name, node.var_attrs.get("__dtype__", "float32"))
                if shp is not None:
                    vals[(id(node), 0)] = torch.empty(shp, dtype=torch_dtype(dt), device="meta")
                    vshape[node.name] = shp
                vdtype[node.name] = dt
                continue
            # fill parameter shapes from the data shape
            _infer_param_shapes(node, vals, vshape, vdtype)
            ins = []
            missing = False
            for inp, idx in node.inputs:
                t = vals.get((id(inp), idx))
                if t is None:
                    missing = True
                    break
                ins.append(t)
            if missing:
                if partial:
                    continue
                unknown = [inp.name for inp, idx in node.inputs if (id(inp), idx) not in vals]
                raise MXNetError("infer_shape: cannot infer shape of inputs %s of node %s" % (unknown, node.name))
            outs = _run_node(node, ins, meta=True)
            for i, o in enumerate(outs):
                if isinstance(o, torch.Tensor) and o.device.type != "meta":
                    o = torch.empty(o.shape, dtype=o.dtype, device="meta")    # constants (zeros / eye / arange) carry no data here
                vals[(id(node), i)] = o
        out_shapes, out_types = [], []
        for n, i in self._outputs:
            t = vals.get((id(n), i))
            out_shapes.append(tuple(t.shape) if t is not None else None)
            out_types.append(np_dtype(t.dtype) if t is not None else None)
        if partial and solved_heads is not None:      # what the bidirectional pass knows about outputs the forward pass could not reach
            out_shapes = [o if o is not None else h for o, h in zip(out_shapes, solved_heads)]
        return vshape, out_shapes, vdtype, out_types

    def infer_shape(self, *args, **kwargs):
        if args:
            kwargs = dict(zip(self.list_arguments(), args))
        try:
            vshape, out_shapes, _, _ = self._infer(kwargs)
        except MXNetError as e:
            if "cannot infer shape" not in str(e):
                raise
            # not enough information: the reference warns and returns (None, None, None) (symbol.py:1001-1015)
            import warnings
            warnings.warn("Cannot decide shape for some arguments: %s" % e, stacklevel=2)
            return None, None, None
        arg_shapes = [vshape.get(n) for n in self.list_arguments()]
        aux_shapes = [vshape.get(n) for n in self.list_auxiliary_states()]
        if any(v is None for v in arg_shapes + aux_shapes + list(out_shapes)):
            return None, None, None
        return arg_shapes, out_shapes, aux_shapes

    def infer_shape_partial(self, *args, **kwargs):
        """Like infer_shape but returns what could be determined; unknown shapes are ``()``."""
        if args:
            kwargs = dict(zip(self.list_arguments(), args))
        vshape, out_shapes, _, _ = self._infer(kwargs, partial=True)
        declared = {n.name: tuple(n.var_attrs["__shape__"]) for n in self._topo() if n.op is None and "__shape__" in n.var_attrs}
        given = {k: tuple(v) for k, v in kwargs.items() if v is not None}
        part = lambda n: vshape.get(n, given.get(n, declared.get(n, ())))      # noqa: E731 -- partly known shapes keep their 0s
        arg_shapes = [part(n) for n in self.list_arguments()]
        aux_shapes = [part(n) for n in self.list_auxiliary_states()]
        return arg_shapes, [s if s is not None else () for s in out_shapes], aux_shapes

    def infer_type(self, *args, **kwargs):
        """Propagate dtypes through the graph: an operator's outputs take the dtype of its first typed input (Cast and
        friends their ``dtype`` attribute); untyped parameter variables take the dtype of the data they are combined with."""
        if args:
            kwargs = dict(zip(self.list_arguments(), args))
        known = {k: np_dtype(torch_dtype(v)) for k, v in kwargs.items() if v is not None}
        ty = {}
        fixed = {"Cast": "dtype", "cast": "dtype", "one_hot": "dtype", "_zeros": "dtype", "_ones": "dtype", "_arange": "dtype",
                 "_eye": "dtype", "_full": "dtype", "topk": "dtype", "argsort": "dtype"}
        order = self._topo()
        for _pass in range(2):          # second pass lets variables typed by a later consumer reach earlier nodes
            for n in order:
                if n.op is None:
                    if n.name in known:
                        ty[(id(n), 0)] = known[n.name]
                    elif "__dtype__" in n.var_attrs and (id(n), 0) not in ty:
                        ty[(id(n), 0)] = np_dtype(torch_dtype(n.var_attrs["__dtype__"]))
                    continue
                in_t = [ty.get((id(i), k)) for i, k in n.inputs]
                first = next((t for t in in_t if t is not None), None)
                if n.op in fixed and n.attrs.get(fixed[n.op]) is not None:
                    out_t = np_dtype(torch_dtype(n.attrs[fixed[n.op]]))
                else:
                    out_t = first
                if first is not None:
                    for (i, k), t in zip(n.inputs, in_t):
                        if t is None and i.op is None:
                            ty[(id(i), k)] = first
                if out_t is not None:
                    for k in range(n.num_outputs):
                        ty[(id(n), k)] = out_t
        args_l = [n for n in order if n.op is None and not n.is_aux]
        auxs_l = [n for n in order if n.op is None and n.is_aux]
        if not known and not any("__dtype__" in n.var_attrs for n in args_l):
            return None, None, None
        arg_types = [ty.get((id(n), 0), np.float32) for n in args_l]
        aux_types = [ty.get((id(n), 0), np.float32) for n in auxs_l]
        out_types = [ty.get((id(n), i), np.float32) for n, i in self._outputs]
        return arg_types, out_types, aux_types

    # ---- execution ---------------------------------------------------------------------------------
    def eval_nd(self, feed, group2ctx=None, monitor=None):
        """Run the graph on NDArrays: feed maps variable name -> NDArray.  Returns list of outputs.
        ``group2ctx``: model parallelism by placement -- a node whose ``ctx_group`` attribute (AttrScope) is a key of
        the map runs on that context; inputs coming from another device are copied there first (the reference inserts
        _CrossDeviceCopy nodes for this: graph_executor.cc:1349-1353, nnvm place_device.cc)."""
        vals = {}
        topo = self._topo()
        # ``monitor(name, array, when)``: called with every operator output (``<node>_output``) right after it is computed
        # and, during backward, with the gradient arriving at every operator input (``<node>_backward_<input>``): the
        # executor's monitor callback of the reference (graph_executor.cc:ExecuteMonCallback).  Fusion is off while monitoring
        # so that every node of the graph shows up.
        deferred, tails = self._fusion_plan(topo) if (not group2ctx and monitor is None and getenv("MXB200_GRAPH_FUSION", True)) \
            else ({}, {})

        def input_names(node):
            names = ["data"] + [x for x, _ in _param_args(node.op, node.attrs)]
            if node.op in _LABEL_OPS:
                names.append(_LABEL_OPS[node.op])
            if len(node.inputs) == 2 and len(names) == 1:
                names = ["lhs", "rhs"]
            kw = node.attrs.get("__kwarg_inputs__")
            if kw:
                names = list(kw)
            return [names[i] if i < len(names) else "arg%d" % i for i in range(len(node.inputs))]

        def run(node):
            ins = [vals[(id(inp), idx)] for inp, idx in node.inputs]
            if group2ctx:
                tgt = group2ctx.get(node.var_attrs.get("__ctx_group__"))
                if tgt is not None:
                    ins = [x.as_in_context(tgt) if hasattr(x, "as_in_context") and x.context != tgt else x for x in ins]
            outs = _run_node(node, ins, meta=False)
            for i, o in enumerate(outs):
                vals[(id(node), i)] = o
            if monitor is not None:
                for i, o in enumerate(outs):
                    monitor(node.name + ("_output" if len(outs) == 1 else "_output%d" % i), o, "forward")
                for x, aname in zip(ins, input_names(node)):
                    monitor("%s_backward_%s" % (node.name, aname), x, "register")
        by_node = feed.get("__by_node__") or {}
        for node in topo:
            if node.op is None:
                if id(node) in by_node:
                    vals[(id(node), 0)] = by_node[id(node)]
                    continue
                if node.name not in feed:
                    raise MXNetError("eval: missing input %s" % node.name)
                vals[(id(node), 0)] = feed[node.name]
                continue
            if id(node) in deferred:
                continue                      # runs (fused or not) when its group's last node is reached
            group = tails.get(id(node))
            if group is not None:
                conv, bn, add, other, relu = group
                get = lambda ins: [vals[(id(inp), idx)] for inp, idx in ins]   # noqa: E731
                fused = _run_fused_group(conv, get(conv.inputs) if conv is not None else None, bn, get(bn.inputs[1:]),
                                         None if conv is not None else get(bn.inputs[:1])[0],
                                         vals[(id(other[0]), other[1])] if other is not None else None, relu)
                if fused is not None:
                    vals[(id(node), 0)] = fused
                    continue
                for member in (conv, bn, add):      # conditions not met at run time (CPU, dtype, width): op by op
                    if member is not None and member is not node:
                        run(member)
            run(node)
        return [vals[(id(n), i)] for n, i in self._outputs]

    def _fusion_plan(self, topo):
        """Pattern pass over the graph (the reference has none; cuDNN had no fused BN+add+ReLU in that era):
            [Convolution 1x1 ->] BatchNorm [-> (broadcast|elemwise)_add] [-> Activation(relu)]
        executes as: tcgen05 GEMM whose epilogue also emits the BatchNorm statistics, then ONE fused apply kernel that reads
        the residual and writes the ReLU output + a 1-bit mask.  A node joins a group only when its value has no other
        consumer and is not a graph output.
        Returns (deferred node ids, {id(last node): (conv | None, bn, add | None, (other, idx) | None, relu?)})."""
        cached = getattr(self, "_fusion_cache", None)
        if cached is not None and cached[0] == len(topo):
            return cached[1], cached[2]
        users = {}
        for n in topo:
            for inp, idx in n.inputs:
                users.setdefault((id(inp), idx), []).append(n)
        for n, i in self._outputs:
            users.setdefault((id(n), i), []).append(None)
        deferred, tails = {}, {}
        adds = ("broadcast_add", "elemwise_add", "_Plus", "_plus", "_add", "broadcast_plus")

        def only_user(n):
            u = users.get((id(n), 0), [])
            return u[0] if len(u) == 1 and u[0] is not None else None

        def is_relu(n):
            return n is not None and n.op == "Activation" and str(n.attrs.get("act_type", "relu")) == "relu"

        def ones(v, n=2):
            v = _tuple_attr(v)
            return v is None or all(int(x) == 1 for x in v)

        def conv_of(bn):
            c, idx = bn.inputs[0]
            a = c.attrs
            if c.op != "Convolution" or idx != 0 or only_user(c) is not bn or id(c) in deferred:
                return None
            k, p = _tuple_attr(a.get("kernel")), _tuple_attr(a.get("pad"))
            if k is None or tuple(int(x) for x in k) != (1, 1) or not (p is None or all(int(x) == 0 for x in p)):
                return None
            if not ones(a.get("dilate")) or int(a.get("num_group", 1)) != 1 or not _truthy(a.get("no_bias", False)):
                return None
            st = _tuple_attr(a.get("stride"))
            if st is not None and len(set(int(x) for x in st)) != 1:
                return None
            return c
        for n in topo:
            if n.op != "BatchNorm" or _truthy(n.attrs.get("output_mean_var", False)) or int(n.attrs.get("axis", 1)) != 1:
                continue
            if any(len(users.get((id(n), i), [])) for i in (1, 2)):
                continue
            conv = conv_of(n)
            u = only_user(n)
            tail, add, other, relu = None, None, None, False
            if is_relu(u):
                tail, relu = u, True
            elif u is not None and u.op in adds and len(u.inputs) == 2 and is_relu(only_user(u)):
                rest = [x for x in u.inputs if not (x[0] is n and x[1] == 0)]
                if len(rest) == 1 and not (rest[0][0].op == "BatchNorm" and id(rest[0][0]) in deferred):
                    tail, add, other, relu = only_user(u), u, rest[0], True
            if tail is None:
                if conv is None:
                    continue
                tail = n                                   # Convolution + BatchNorm only (e.g. the projection shortcut)
            for member in (conv, n, add):
                if member is not None and member is not tail:
                    deferred[id(member)] = True
            tails[id(tail)] = (conv, n, add, other, relu)
        self._fusion_cache = (len(topo), deferred, tails)
        return deferred, tails


    def eval(self, ctx=None, **kwargs):
        return self.eval_nd(kwargs)

    def simple_bind(self, ctx, grad_req="write", type_dict=None, stype_dict=None, group2ctx=None, shared_arg_names=None,
                    shared_exec=None, shared_buffer=None, **kwargs):
        from ..executor import Executor
        return Executor.simple_bind(self, ctx, grad_req, type_dict, group2ctx, kwargs)

    def bind(self, ctx, args, args_grad=None, grad_req="write", aux_states=None, group2ctx=None, shared_exec=None):
        from ..executor import Executor
        return Executor(self, ctx, args, args_grad, grad_req, aux_states, group2ctx)

    # ---- serialization --------------------------------------------------------------------------------
    def tojson(self):
        order = self._topo()
        index = {id(n): i for i, n in enumerate(order)}
        nodes, arg_nodes, row_ptr = [], [], [0]
        for i, n in enumerate(order):
            d = {"op": "null" if n.op is None else n.op, "name": n.name,
                 "inputs": [[index[id(inp)], idx, 0] for inp, idx in n.inputs]}
            attrs = {k: _attr_str(v) for k, v in n.attrs.items()}
            attrs.update({k: _attr_str(v) for k, v in n.var_attrs.items()})
            if n.op is None and n.is_aux:
                attrs["__is_aux__"] = "1"
            if attrs:
                d["attrs"] = attrs
            nodes.append(d)
            if n.op is None:
                arg_nodes.append(i)
            row_ptr.append(row_ptr[-1] + n.num_outputs)
        heads = [[index[id(n)], i, 0] for n, i in self._outputs]
        return json.dumps({"nodes": nodes, "arg_nodes": arg_nodes, "node_row_ptr": row_ptr, "heads": heads,
                           "attrs": {"mxnet_version": ["int", 10301], "framework": ["str", "mxb200"]}}, indent=2)

    def save(self, fname):
        with open(fname, "w") as f:
            f.write(self.tojson())

    def __deepcopy__(self, memo):
        return load_json(self.tojson())

    def __getstate__(self):
        return {"json": self.tojson()}

    def __setstate__(self, state):
        self._outputs = load_json(state["json"])._outputs


def _tuple_attr(v):
    if v is None:
        return None
    if isinstance(v, str):
        v = [x for x in v.strip("()[] ").split(",") if x.strip()]
    if isinstance(v, (int, float)):
        return (int(v), int(v))
    return tuple(int(float(x)) for x in v)


def _attr_str(v):
    if isinstance(v, (tuple, list)):
        return "(" + ", ".join(str(x) for x in v) + ("," if len(v) == 1 else "") + ")"
    return str(v)


def _parse_attr(s):
    if not isinstance(s, str):
        return s
    t = s.strip()
    if t in ("True", "true"):
        return True
    if t in ("False", "false"):
        return False
    if t == "None":
        return None
    if t.startswith("(") or t.startswith("["):
        inner = t[1:-1].strip()
        if not inner:
            return ()
        return tuple(_parse_attr(x) for x in inner.split(",") if x.strip() != "")
    try:
        return int(t)
    except ValueError:
        pass
    try:
        return float(t)
    except ValueError:
        return s


def load_json(json_str):
    g = json.loads(json_str)
    nodes = []
    for d in g["nodes"]:
        attrs = d.get("attrs", d.get("attr", d.get("param", {}))) or {}
        if d["op"] == "null":
            n = _Node(None, d["name"])
            n.is_aux = attrs.get("__is_aux__") == "1"
            n.var_attrs = {k: _parse_attr(v) if k in ("__shape__",) else v for k, v in attrs.items() if k != "__is_aux__"}
        else:
            n = _Node(d["op"], d["name"], {k: _parse_attr(v) for k, v in attrs.items() if not k.startswith("__")})
            n.var_attrs = {k: v for k, v in attrs.items() if k.startswith("__")}
        nodes.append(n)
    for n, d in zip(nodes, g["nodes"]):
        n.inputs = [(nodes[i[0]], i[1]) for i in d["inputs"]]
        if n.op is not None:
            n.num_outputs = _num_outputs(n.op, n.attrs)
        # legacy graphs: aux inputs of BatchNorm are variables named *moving_*
        if n.op == "BatchNorm":
            for inp, _ in n.inputs[3:5]:
                inp.is_aux = True
    return Symbol([(nodes[h[0]], h[1]) for h in g["heads"]])


def load(fname):
    with open(fname) as f:
        return load_json(f.read())


def var(name, attr=None, shape=None, lr_mult=None, wd_mult=None, dtype=None, init=None, stype=None, **kwargs):
    if not isinstance(name, str):
        raise TypeError("Expect a string for variable `name`")
    n = _Node(None, name)
    a = dict(attr or {})
    if shape is not None:
        a["__shape__"] = (int(shape),) if isinstance(shape, (int, np.integer)) else tuple(shape)
    if lr_mult is not None:
        a["__lr_mult__"] = lr_mult
    if wd_mult is not None:
        a["__wd_mult__"] = wd_mult
    if dtype is not None:
        a["__dtype__"] = dtype_name(dtype)
    if init is not None:
        a["__init__"] = init if isinstance(init, str) else init.dumps()
    if stype is not None:
        a["__storage_type__"] = str({"default": 0, "row_sparse": 1, "csr": 2}[stype])    # _STORAGE_TYPE_STR_TO_ID
    for k, v in kwargs.items():
        if k.startswith("__") and k.endswith("__"):
            a[k] = str(v)
        else:
            raise ValueError("Attribute name=%s is not supported. Additional attributes must start and end with double "
                             "underscores, e.g, __yourattr__" % k)
    final = {}
    for k, v in a.items():
        _put_attr(final, k, v)
    from ..attribute import current_attrs
    for k, v in current_attrs().items():        # e.g. AttrScope(ctx_group='dev1') -> ctx_group + __ctx_group__
        _put_attr(final, k, v, override=False)
    n.var_attrs = final
    return Symbol([(n, 0)])


Variable = var


def Group(symbols):
    outs = []
    for s in symbols:
        outs.extend(s._outputs)
    return Symbol(outs)


def _shp(shape):
    return (int(shape),) if isinstance(shape, (int, np.integer)) else tuple(int(x) for x in shape)


def zeros(shape, dtype=None, **kwargs):
    return _make_op("_zeros", [], {"shape": _shp(shape), "dtype": dtype_name(dtype or "float32")}, name=kwargs.get("name"))


def ones(shape, dtype=None, **kwargs):
    return _make_op("_ones", [], {"shape": _shp(shape), "dtype": dtype_name(dtype or "float32")}, name=kwargs.get("name"))


def full(shape, val, dtype=None, **kwargs):
    return ones(shape, dtype=dtype, **kwargs) * val


def arange(start, stop=None, step=1.0, repeat=1, name=None, dtype=None):
    return _create("_arange", [], dict(start=start, stop=stop, step=step, repeat=repeat, dtype=dtype_name(dtype or "float32"),
                                       name=name))


def eye(N, M=0, k=0, dtype=None, **kwargs):
    return _create("_eye", [], dict(N=N, M=M, k=k, dtype=dtype_name(dtype or "float32"), **kwargs))


def _sym_or_scalar(name, sym_op, scalar_op, rscalar_op, py_fn):
    """Module-level binary helpers with MXNet's dispatch: Symbol x Symbol -> broadcast op, Symbol x number -> scalar op,
    number x number -> plain python (symbol.py:pow / maximum / minimum / hypot of the reference)."""
    def f(left, right):
        ls, rs = isinstance(left, Symbol), isinstance(right, Symbol)
        if ls and rs:
            return _make_op(sym_op, [left, right], {})
        if ls:
            return _make_op(scalar_op, [left], {"scalar": float(right)})
        if rs:
            return _make_op(rscalar_op, [right], {"scalar": float(left)})
        if isinstance(left, (int, float)) and isinstance(right, (int, float)):
            return py_fn(left, right)
        raise TypeError("types (%s, %s) not supported" % (str(type(left)), str(type(right))))
    f.__name__ = name
    return f


pow = _sym_or_scalar("pow", "broadcast_power", "_power_scalar", "_rpower_scalar", lambda a, b: a ** b)   # noqa: A001
power = pow
maximum = _sym_or_scalar("maximum", "broadcast_maximum", "_maximum_scalar", "_maximum_scalar", lambda a, b: a if a > b else b)
minimum = _sym_or_scalar("minimum", "broadcast_minimum", "_minimum_scalar", "_minimum_scalar", lambda a, b: a if a < b else b)
hypot = _sym_or_scalar("hypot", "broadcast_hypot", "_hypot_scalar", "_hypot_scalar", lambda a, b: (a * a + b * b) ** 0.5)


def _make_op(op, inputs, attrs, name=None, hint=None):
    hint = hint or op.lower().lstrip("_")
    name = _NameManager.resolve(name, hint)
    nout = _num_outputs(op, attrs)
    node = _Node(op, name, attrs, [s._outputs[0] for s in inputs], nout)
    from ..attribute import current_attrs
    for k, v in current_attrs().items():
        _put_attr(node.var_attrs, k, v)
    return Symbol([(node, i) for i in range(nout)]) if nout > 1 else Symbol([(node, 0)])


def _create(op, args, kwargs):
    """mx.sym.<op>(*args, **kwargs): Symbol inputs become graph edges, the rest become attributes; parameter
    variables (weight/bias/gamma/...) and label variables are auto-created with the reference naming."""
    name = kwargs.pop("name", None)
    kwargs.pop("out", None)                     # generated symbol functions accept and ignore ``out`` (symbol/register.py)
    user_attr = dict(kwargs.pop("attr", None) or {})
    for v in user_attr.values():
        if not isinstance(v, str):
            raise ValueError("Attributes need to be string")
    for k in _HIDDEN_ATTR_KEYS:                      # e.g. Convolution(..., lr_mult=1): node attribute, not an operator argument
        if k in kwargs and not isinstance(kwargs[k], Symbol):
            user_attr[k] = str(kwargs.pop(k))
    for k in [k for k in kwargs if k.startswith("__") and k.endswith("__") and not isinstance(kwargs[k], Symbol)]:
        user_attr[k] = str(kwargs.pop(k))            # e.g. Activation(..., __layout__='NCHW')
    from ..attribute import current_attrs
    node_attrs = {}
    for k, v in user_attr.items():
        _put_attr(node_attrs, k, v)
    for k, v in current_attrs().items():
        _put_attr(node_attrs, k, v, override=False)
    inherit = {k: v for k, v in node_attrs.items() if k.startswith("__") and k.endswith("__")}   # reach the auto-created variables
    sym_kwargs = {k: v for k, v in kwargs.items() if isinstance(v, Symbol)}
    attrs = {k: v for k, v in kwargs.items() if not isinstance(v, Symbol) and v is not None}
    if op in ("topk", "sort", "argsort") and "axis" in kwargs and kwargs["axis"] is None:
        attrs["axis"] = "None"               # an explicit None (= flatten) differs from these operators' default axis -1
    hint = op.lower().lstrip("_")
    node_name = _NameManager.resolve(name, hint)
    inputs = [a for a in args if isinstance(a, Symbol)]
    pos_attrs = [a for a in args if not isinstance(a, Symbol) and a is not None]
    if pos_attrs:
        # positional non-symbol args (e.g. nd-style clip(x, a_min, a_max)) -> named by the nd op signature
        import inspect
        fn = _ndop.get(op)
        if fn is not None:
            params = [p for p in inspect.signature(fn).parameters.values() if p.kind in (p.POSITIONAL_OR_KEYWORD,)]
            names = [p.name for p in params]
            k = 0
            for a, pname in zip(args, names):
                if not isinstance(a, Symbol) and a is not None:
                    attrs[pname] = a
    order = []
    if _has_param_args(op) or op in _LABEL_OPS:
        spec = [("data", False)] + _param_args(op, attrs)
        if op in _LABEL_OPS:
            spec = spec + [(_LABEL_OPS[op], False)]
        provided = list(inputs)
        for i, (aname, is_aux) in enumerate(spec):
            if aname in sym_kwargs:
                s = sym_kwargs[aname]
            elif provided:
                s = provided.pop(0)
            else:
                if aname == "bias" and attrs.get("no_bias", False):
                    continue
                if aname == "gamma" and op == "LeakyReLU" and attrs.get("act_type", "leaky") != "prelu":
                    continue
                # inputs that are not given become free variables "<node>_<argument>" (to be composed later)
                s = var("%s_%s" % (node_name, aname))
                for k, v in inherit.items():
                    s._outputs[0][0].var_attrs.setdefault(k, v)
                if op == "LeakyReLU" and aname == "gamma":
                    # the PReLU slope starts at 0.25 whatever the global initializer is (leaky_relu.cc FSetInputVarAttrOnCompose)
                    s._outputs[0][0].var_attrs.setdefault("__init__", '["constant", {"value": 0.25}]')
                if is_aux:
                    s._outputs[0][0].is_aux = True
            if is_aux:
                s._outputs[0][0].is_aux = True
            order.append(s)
    elif op == "Custom":
        # inputs = the prop's arguments followed by its auxiliary states (custom.cc ListArguments); whatever is not given
        # becomes a free variable "<node>_<name>", auxiliary ones flagged as such
        prop = _custom_prop(attrs)
        provided = list(inputs)
        for aname in list(prop.list_arguments()) + list(prop.list_auxiliary_states()):
            is_aux = aname in prop.list_auxiliary_states()
            if aname in sym_kwargs:
                s_ = sym_kwargs[aname]
            elif provided:
                s_ = provided.pop(0)
            else:
                s_ = var("%s_%s" % (node_name, aname))
            if is_aux and s_._outputs[0][0].op is None:
                s_._outputs[0][0].is_aux = True
            order.append(s_)
    else:
        order = inputs + [sym_kwargs[k] for k in sym_kwargs]
        if sym_kwargs:
            attrs["__kwarg_inputs__"] = tuple(sym_kwargs.keys())
        if not order and not op.startswith(("_random_", "_zeros", "_ones", "_full", "_arange", "_eye")):
            # no inputs given: every required tensor argument becomes a free variable "<node>_<argument>"
            import inspect
            fn = _ndop.get(op)
            try:
                params = list(inspect.signature(fn).parameters.values()) if fn is not None else []
            except (TypeError, ValueError):
                params = []
            for prm in params:
                if prm.kind not in (prm.POSITIONAL_ONLY, prm.POSITIONAL_OR_KEYWORD) or prm.default is not inspect.Parameter.empty \
                        or prm.name in attrs:
                    break
                order.append(var("%s_%s" % (node_name, prm.name)))
    nout = _num_outputs(op, attrs)
    node = _Node(op, node_name, attrs, [s._outputs[0] for s in order], nout)
    node.var_attrs.update(node_attrs)
    return Symbol([(node, i) for i in range(nout)]) if nout > 1 else Symbol([(node, 0)])


_SCALAR_OPS = {
    "_plus_scalar": lambda x, s, rev: x + s, "_minus_scalar": lambda x, s, rev: (s - x) if rev else (x - s),
    "_mul_scalar": lambda x, s, rev: x * s, "_div_scalar": lambda x, s, rev: (s / x) if rev else (x / s),
    "_power_scalar": lambda x, s, rev: (s ** x) if rev else (x ** s),
    "_greater_scalar": lambda x, s, rev: x > s, "_lesser_scalar": lambda x, s, rev: x < s,
    "_equal_scalar": lambda x, s, rev: x == s,
}


def _run_fused_group(conv, conv_ins, bn, bn_params, x_in, other, relu):
    """[1x1 Convolution +] BatchNorm [+ residual] [+ ReLU] on the fused kernels, or None when they do not apply to these
    tensors (the caller then runs the member nodes one by one)."""
    gamma, beta, mm, mv = bn_params
    data = conv_ins[0] if conv is not None else x_in
    x = data._t
    if not (x.is_cuda and x.dim() == 4):
        return None
    from ..ops import functional as K
    from ..ndarray.ndarray import _op
    attrs = bn.attrs
    training = autograd.is_training() and not _truthy(attrs.get("use_global_stats", False))
    eps, mom = float(attrs.get("eps", 1e-3)), float(attrs.get("momentum", 0.9))
    fix_gamma = _truthy(attrs.get("fix_gamma", True))
    stride = 1
    if conv is not None:
        st = _tuple_attr(conv.attrs.get("stride"))
        stride = int(st[0]) if st else 1

    class _NoFuse(Exception):
        pass

    def g(xt, *ts):
        ts = list(ts)
        w = ts.pop(0) if conv is not None else None
        gm, bt = ts[0], ts[1]
        res = ts[2] if len(ts) > 2 else None
        if w is None:
            y = K.batch_norm_nchw(xt, gm, bt, mm._t, mv._t, eps, mom, training, fix_gamma, residual=res, relu=relu)
        else:
            y = K.conv1x1_bn_nchw(xt, w, stride, gm, bt, mm._t, mv._t, eps, mom, training, fix_gamma, residual=res, relu=relu)
        if y is None:
            raise _NoFuse()
        return y
    args = [data] + ([conv_ins[1]] if conv is not None else []) + [gamma, beta] + ([other] if other is not None else [])
    try:
        return _op(g, *args)
    except _NoFuse:
        return None


def _run_node(node, ins, meta):
    """Execute one node.  meta=True: ins are meta tensors (shape inference); else NDArrays."""
    op, attrs = node.op, dict(node.attrs)
    kw_inputs = attrs.pop("__kwarg_inputs__", None)
    if meta:
        ins = [NDArray(t, cpu()) for t in ins]
    if op in _SCALAR_OPS:
        out = _SCALAR_OPS[op](ins[0], attrs["scalar"], attrs.get("reverse", False))
    elif op in ("_zeros", "_ones"):
        if meta:   # shape inference: stay on the meta device
            import torch as _torch
            from ..base import torch_dtype as _td
            out = NDArray(_torch.empty(tuple(attrs["shape"]), dtype=_td(attrs.get("dtype") or "float32"), device="meta"), cpu())
        else:
            out = (nd.zeros if op == "_zeros" else nd.ones)(attrs["shape"], dtype=attrs.get("dtype"))
    else:
        fn = _ndop.get(op)
        if fn is None:
            raise MXNetError("operator %s is not registered" % op)
        if _has_param_args(op) or op in _LABEL_OPS:
            spec = ["data"] + [a for a, _ in _param_args(op, attrs)]
            if op in _LABEL_OPS:
                spec.append(_LABEL_OPS[op])
            present = []
            for aname in spec:
                if aname == "bias" and attrs.get("no_bias", False):
                    continue
                if aname == "gamma" and op == "LeakyReLU" and attrs.get("act_type", "leaky") != "prelu":
                    continue
                present.append(aname)
            kwargs = dict(zip(present, ins))
            data = kwargs.pop("data")
            if meta and op in _LABEL_OPS:
                kwargs.pop(_LABEL_OPS[op], None)
                out = data if op not in ("SoftmaxOutput", "Softmax") else data
            else:
                out = fn(data, **kwargs, **attrs)
        elif kw_inputs:
            npos = len(ins) - len(kw_inputs)
            out = fn(*ins[:npos], **dict(zip(kw_inputs, ins[npos:])), **attrs)
        else:
            out = fn(*ins, **attrs)
    outs = list(out) if isinstance(out, (list, tuple)) else [out]
    if meta:
        return [o._t for o in outs]
    return outs


def _infer_param_shapes(node, vals, vshape, vdtype):
    """Fill shapes of parameter/label variables feeding ``node`` from its data input (infer_shape rules)."""
    op = node.op
    if op == "Custom":
        prop = _custom_prop(node.attrs)
        nargs = len(prop.list_arguments())
        known = [vals.get((id(n), i)) for n, i in node.inputs]
        if any(k is None for k in known) and all(k is not None for k in known[:nargs]) and nargs:
            ishapes, _o, aux_shapes = prop.infer_shape([list(k.shape) for k in known[:nargs]])
            full = list(ishapes) + list(aux_shapes)
            for (n, i), k, shp in zip(node.inputs, known, full):
                if k is None and n.op is None:
                    vals[(id(n), i)] = torch.empty(tuple(shp), dtype=known[0].dtype, device="meta")
                    vshape[n.name] = tuple(shp)
                    vdtype[n.name] = dtype_name(known[0].dtype)
        return
    if op == "RNN":
        # flat cuDNN-layout parameter vector of the fused RNN op: size from (input size, state size, layers, directions, mode)
        a = node.attrs
        names = list(a.get("__kwarg_inputs__", ())) or ["data", "parameters", "state", "state_cell"][:len(node.inputs)]
        byname = dict(zip(names, node.inputs))
        if "data" in byname and "parameters" in byname:
            d = vals.get((id(byname["data"][0]), byname["data"][1]))
            pn, pidx = byname["parameters"]
            if d is not None and pn.op is None and (id(pn), pidx) not in vals:
                g = {"rnn_relu": 1, "rnn_tanh": 1, "lstm": 4, "gru": 3}[str(a.get("mode", "lstm"))]
                H, L = int(a["state_size"]), int(a.get("num_layers", 1))
                nd_ = 2 if _truthy(a.get("bidirectional", False)) else 1
                size, inp = 0, int(d.shape[2])
                for _l in range(L):
                    size += nd_ * (g * H * (inp + H) + 2 * g * H)
                    inp = H * nd_
                dt = dtype_name(d.dtype)
                vals[(id(pn), pidx)] = torch.empty((size,), dtype=d.dtype, device="meta")
                vshape[pn.name] = (size,)
                vdtype[pn.name] = dt
        if "data" in byname:
            d = vals.get((id(byname["data"][0]), byname["data"][1]))
            if d is not None:
                nd_ = 2 if _truthy(a.get("bidirectional", False)) else 1
                sshape = (int(a.get("num_layers", 1)) * nd_, int(d.shape[1]), int(a["state_size"]))
                for sname in ("state", "state_cell"):          # initial states: (layers * directions, batch, state size)
                    if sname in byname:
                        sn, sidx = byname[sname]
                        if sn.op is None and (id(sn), sidx) not in vals:
                            vals[(id(sn), sidx)] = torch.empty(sshape, dtype=d.dtype, device="meta")
                            vshape[sn.name] = sshape
                            vdtype[sn.name] = dtype_name(d.dtype)
        return
    if not _has_param_args(op) and op not in _LABEL_OPS:
        return
    a = node.attrs
    data = vals.get((id(node.inputs[0][0]), node.inputs[0][1])) if node.inputs else None
    if data is None:
        return
    ds = tuple(data.shape)
    spec = ["data"] + [x for x, _ in _param_args(op, a)]
    if op in _LABEL_OPS:
        spec.append(_LABEL_OPS[op])
    fused_shapes = {n: (shp, kind) for n, shp, kind, _aux in _fused_resnet_args(a)[1]} if op == "_contrib_FusedResNet" else None
    present = []
    for aname in spec:
        if aname == "bias" and a.get("no_bias", False):
            continue
        if aname == "gamma" and op == "LeakyReLU" and a.get("act_type", "leaky") != "prelu":
            continue
        present.append(aname)

    def rule(aname):
        if fused_shapes is not None:
            return fused_shapes[aname][0] if aname in fused_shapes else None
        if op == "FullyConnected":
            nh = int(a["num_hidden"])
            inn = int(np.prod(ds[1:])) if a.get("flatten", True) else ds[-1]
            return {"weight": (nh, inn), "bias": (nh,)}[aname]
        if op in ("Convolution", "Convolution_v1"):
            k = tuple(a["kernel"]) if not isinstance(a["kernel"], int) else (a["kernel"],) * (len(ds) - 2)
            nf, g = int(a["num_filter"]), int(a.get("num_group", 1))
            return {"weight": (nf, ds[1] // g) + k, "bias": (nf,)}[aname]
        if op == "_contrib_DeformableConvolution":
            if aname == "offset":
                return None
            k = tuple(a["kernel"]) if not isinstance(a["kernel"], int) else (a["kernel"],) * (len(ds) - 2)
            nf, g = int(a["num_filter"]), int(a.get("num_group", 1))
            return {"weight": (nf, ds[1] // g) + k, "bias": (nf,)}[aname]
        if op == "IdentityAttachKLSparseReg":
            return (int(np.prod(ds[1:])),)
        if op == "Deconvolution":
            k = tuple(a["kernel"]) if not isinstance(a["kernel"], int) else (a["kernel"],) * (len(ds) - 2)
            nf, g = int(a["num_filter"]), int(a.get("num_group", 1))
            return {"weight": (ds[1], nf // g) + k, "bias": (nf,)}[aname]
        if op in ("BatchNorm", "InstanceNorm", "BatchNorm_v1", "CuDNNBatchNorm", "_contrib_SyncBatchNorm"):
            return (ds[int(a.get("axis", 1))],)
        if op == "LayerNorm":
            return (ds[int(a.get("axis", -1))],)
        if op in ("Embedding", "_contrib_SparseEmbedding"):
            return (int(a["input_dim"]), int(a["output_dim"]))
        if op == "LeakyReLU":
            return (ds[1],)
        if aname == "label":
            if op in ("SoftmaxOutput", "Softmax"):
                if a.get("multi_output", False):
                    return (ds[0],) + ds[2:]
                return (ds[0],)
            return ds
        return None
    for (inp, idx), aname in zip(node.inputs, present):
        if inp.op is None and (id(inp), idx) in vals and aname not in ("data", "label") and fused_shapes is None:
            # a shape given by the caller must agree with what the operator implies (InferShape consistency check)
            want = rule(aname)
            have = tuple(vals[(id(inp), idx)].shape)
            if want is not None and tuple(want) != have and op != "LeakyReLU":     # (prelu's gamma may be any broadcastable shape)
                raise MXNetError("infer_shape: shape inconsistent for %s of node %s: provided %s, inferred %s"
                                 % (inp.name, node.name, str(have), str(tuple(want))))
        if inp.op is None and (id(inp), idx) not in vals and aname != "data":
            shp = rule(aname)
            if shp is not None:
                dt = vdtype.get(inp.name, inp.var_attrs.get("__dtype__", "float32"))
                if fused_shapes is not None:
                    dt = "float32" if fused_shapes[aname][1] == "f" else dtype_name(data.dtype)
                elif aname not in ("label",) and op not in ("BatchNorm",):
                    dt = dtype_name(data.dtype) if inp.var_attrs.get("__dtype__") is None else dt
                vals[(id(inp), idx)] = torch.empty(shp, dtype=torch_dtype(dt), device="meta")
                vshape[inp.name] = tuple(shp)
                vdtype[inp.name] = dt
