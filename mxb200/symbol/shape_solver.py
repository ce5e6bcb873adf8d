"""Bidirectional shape inference with partially known shapes (reference: NNVM's InferShape pass,
3rdparty/tvm/nnvm/src/pass/infer_shape_type.cc: forward AND backward sweeps over per-operator FInferShape functions
until a fixed point; a dimension 0 means "unknown").

The executor's shape inference proper is a forward evaluation on meta tensors (symbol.py::_infer).  This solver runs in
front of it when some variable shape is missing or has 0-dimensions: every node output gets a list of dimensions (0 =
unknown), per-operator rules unify dimensions between inputs and outputs in both directions, and the variables that
became fully known are handed to the forward pass as if the caller had supplied them.  Typical uses in MXNet code:
``mx.sym.Variable('state', shape=(0, hidden))`` (batch size decided later by the data), weights derived from other
weights, labels inferred from outputs.

Rules exist for the operators where the reference's tests and model code rely on backward inference: same-shape
element-wise arithmetic (the python operators on symbols), shape-preserving unary operators, FullyConnected,
Convolution, SliceChannel, Concat, Flatten, ``_identity_with_attr_like_rhs``, BatchNorm-like normalisation layers,
Dropout, Embedding, SoftmaxOutput-style heads.  Nodes without a rule simply do not propagate."""
import numpy as np

from ..base import MXNetError

_SAME_SHAPE_BINARY = {"elemwise_add", "elemwise_sub", "elemwise_mul", "elemwise_div", "_plus", "_minus", "_mul", "_div", "_Plus", "_Minus",
                      "_Mul", "_Div", "_add", "_sub", "_grad_add", "_maximum", "_minimum", "_power", "_hypot"}
_BROADCAST_FROM_OPERATOR = {"broadcast_add", "broadcast_sub", "broadcast_mul", "broadcast_div"}
_SHAPE_PRESERVING = {"Activation", "relu", "sigmoid", "tanh", "softsign", "LeakyReLU", "Dropout", "BatchNorm", "BatchNorm_v1", "InstanceNorm",
                     "LayerNorm", "CuDNNBatchNorm", "_contrib_SyncBatchNorm", "Cast", "cast", "identity", "_copy", "BlockGrad", "stop_gradient",
                     "negative", "abs", "exp", "log", "sqrt", "square", "softmax", "log_softmax", "SoftmaxActivation", "SoftmaxOutput", "Softmax",
                     "LinearRegressionOutput", "LogisticRegressionOutput", "MAERegressionOutput", "clip", "_plus_scalar", "_minus_scalar",
                     "_mul_scalar", "_div_scalar", "_rminus_scalar", "_rdiv_scalar", "_power_scalar", "zeros_like", "ones_like", "make_loss",
                     "MakeLoss", "L2Normalization", "erf", "sin", "cos", "round", "floor", "ceil", "sign", "reciprocal", "rsqrt", "hard_sigmoid"}


def _truthy(v):
    return v in (True, 1, "True", "true", "1")


def _tuple(v, n):
    if v is None:
        return None
    if isinstance(v, (int, np.integer)):
        return (int(v),) * n
    if isinstance(v, str):
        v = [int(x) for x in v.strip("()[] ").split(",") if x.strip()]
    return tuple(int(x) for x in v)


class _Solver(object):
    def __init__(self):
        self.S = {}
        self.changed = False

    def get(self, key):
        return self.S.get(key)

    def put(self, key, dims, where=""):
        """Unify ``dims`` (0 = unknown) into the shape of ``key``."""
        if dims is None:
            return
        dims = [int(d) for d in dims]
        cur = self.S.get(key)
        if cur is None:
            self.S[key] = dims
            self.changed = True
            return
        if len(cur) != len(dims):
            raise MXNetError("infer_shape: shape inconsistent %s: %s vs %s" % (where, tuple(cur), tuple(dims)))
        for i, (a, b) in enumerate(zip(cur, dims)):
            if a == 0 and b != 0:
                cur[i] = b
                self.changed = True
            elif a != 0 and b != 0 and a != b:
                raise MXNetError("infer_shape: shape inconsistent %s: %s vs %s" % (where, tuple(cur), tuple(dims)))

    def unify(self, keys, where=""):
        """All keys have the same shape."""
        known = [self.S[k] for k in keys if self.S.get(k) is not None]
        if not known:
            return
        ranks = set(len(k) for k in known)
        if len(ranks) != 1:
            raise MXNetError("infer_shape: shape inconsistent %s: ranks %s" % (where, sorted(ranks)))
        merged = [0] * len(known[0])
        for k in known:
            for i, d in enumerate(k):
                if d != 0:
                    if merged[i] != 0 and merged[i] != d:
                        raise MXNetError("infer_shape: shape inconsistent %s: %s" % (where, [tuple(x) for x in known]))
                    merged[i] = d
        for k in keys:
            self.put(k, merged, where)


def _apply(node, s):
    """One node's rule: unify whatever is known among its inputs / outputs."""
    op, a = node.op, node.attrs
    if any(s.get((id(n), i)) is not None and len(s.get((id(n), i))) == 0 for n, i in node.inputs):
        return                      # a scalar-shaped / unknown-rank placeholder: nothing to unify
    ins = [(id(n), i) for n, i in node.inputs]
    out0 = (id(node), 0)
    where = "at node %s (%s)" % (node.name, op)
    if op in _SAME_SHAPE_BINARY or (op in _BROADCAST_FROM_OPERATOR and a.get("__same_shape__")):
        s.unify(ins + [out0], where)
    elif op == "_identity_with_attr_like_rhs":
        s.unify(ins + [out0], where)
    elif op in _SHAPE_PRESERVING:
        s.unify([ins[0], out0], where)      # (labels are left to the forward rules: a (N,) label is legal for (N, 1) data)
    elif op == "FullyConnected":
        nh = int(a["num_hidden"])
        flat = _truthy(a.get("flatten", True))
        d, o = s.get(ins[0]), s.get(out0)
        if d is not None:
            if flat:
                s.put(out0, [d[0], nh], where)
                inn = int(np.prod(d[1:])) if all(x != 0 for x in d[1:]) else 0
            else:
                s.put(out0, list(d[:-1]) + [nh], where)
                inn = d[-1]
            if len(ins) > 1:
                s.put(ins[1], [nh, inn], where)
        if o is not None and d is not None and flat:
            s.put(ins[0], [o[0]] + [0] * (len(d) - 1), where)
        if len(ins) > 1:
            w = s.get(ins[1])
            if w is not None and d is not None and flat and len(d) == 2:
                s.put(ins[0], [0, w[1]], where)
            if w is None:
                s.put(ins[1], [nh, 0], where)
        if len(ins) > 2 and not _truthy(a.get("no_bias", False)):
            s.put(ins[2], [nh], where)
    elif op in ("Convolution", "Convolution_v1"):
        d, o = s.get(ins[0]), s.get(out0)
        rank = len(d) if d is not None else (len(o) if o is not None else None)
        if rank is None or rank < 3:
            return
        nsp = rank - 2
        k = _tuple(a["kernel"], nsp)
        st = _tuple(a.get("stride"), nsp) or (1,) * nsp
        dl = _tuple(a.get("dilate"), nsp) or (1,) * nsp
        pd = _tuple(a.get("pad"), nsp) or (0,) * nsp
        nf = int(a["num_filter"])
        if d is not None:
            sp = [(d[2 + i] + 2 * pd[i] - dl[i] * (k[i] - 1) - 1) // st[i] + 1 if d[2 + i] != 0 else 0 for i in range(nsp)]
            s.put(out0, [d[0], nf] + sp, where)
        if o is not None:
            # stride 1 inverts exactly; strided convolutions keep the smallest input size that produces the output
            sp = [(o[2 + i] - 1) * st[i] - 2 * pd[i] + dl[i] * (k[i] - 1) + 1 if o[2 + i] != 0 else 0 for i in range(nsp)]
            s.put(ins[0], [o[0], 0] + (sp if all(x == 1 for x in st) else [0] * nsp), where)
    elif op in ("SliceChannel", "split"):
        n = int(a["num_outputs"])
        axis = int(a.get("axis", 1))
        squeeze = _truthy(a.get("squeeze_axis", False))
        d = s.get(ins[0])
        outs = [(id(node), i) for i in range(n)]
        if d is not None:
            ax = axis % len(d)
            part = d[ax] // n if d[ax] != 0 else 0
            od = [x for i, x in enumerate(d) if not (squeeze and i == ax)] if squeeze else [part if i == ax else x for i, x in enumerate(d)]
            for ok in outs:
                s.put(ok, od, where)
        known = [s.get(ok) for ok in outs if s.get(ok) is not None]
        if known and d is not None:
            ax = axis % len(d)
            o = known[0]
            back = list(o[:ax]) + [d[ax]] + list(o[ax:]) if squeeze else [d[ax] if i == ax else x for i, x in enumerate(o)]
            if not squeeze and o[ax] != 0:
                back[ax] = o[ax] * n
            s.put(ins[0], back, where)
        if known:
            s.unify(outs, where)
    elif op in ("Concat", "concat"):
        dim = int(a.get("dim", 1))
        shapes = [s.get(k) for k in ins]
        o = s.get(out0)
        ref = next((x for x in shapes + [o] if x is not None), None)
        if ref is None:
            return
        ax = dim % len(ref)
        other = [0] * len(ref)
        for x in shapes + [o]:
            if x is None:
                continue
            for i, v in enumerate(x):
                if i != ax and v != 0:
                    if other[i] not in (0, v):
                        raise MXNetError("infer_shape: shape inconsistent %s" % where)
                    other[i] = v
        for k, x in zip(ins, shapes):
            s.put(k, [x[ax] if (x is not None and i == ax) else other[i] for i in range(len(ref))], where)
        shapes = [s.get(k) for k in ins]
        tot = sum(x[ax] for x in shapes) if all(x is not None and x[ax] != 0 for x in shapes) else 0
        s.put(out0, [tot if i == ax else other[i] for i in range(len(ref))], where)
        o = s.get(out0)
        if o is not None and o[ax] != 0:
            unknown = [k for k, x in zip(ins, shapes) if x is None or x[ax] == 0]
            if len(unknown) == 1:
                rest = sum(x[ax] for x in shapes if x is not None and x[ax] != 0)
                cur = s.get(unknown[0])
                s.put(unknown[0], [o[ax] - rest if i == ax else cur[i] for i in range(len(ref))], where)
    elif op in ("Flatten", "flatten"):
        d = s.get(ins[0])
        if d is not None:
            s.put(out0, [d[0], int(np.prod(d[1:])) if all(x != 0 for x in d[1:]) else 0], where)
        o = s.get(out0)
        if o is not None and d is not None:
            s.put(ins[0], [o[0]] + [0] * (len(d) - 1), where)
    elif op in ("_ones", "_zeros", "_full", "ones", "zeros", "full") and not ins:
        shp = _tuple(a.get("shape"), 1)
        if shp is not None and all(x > 0 for x in shp):
            s.put(out0, list(shp), where)
    elif op in ("Reshape", "reshape"):
        # backward only: a fully known output fixes the one missing input dimension (element count is preserved)
        d, o = s.get(ins[0]), s.get(out0)
        if d is not None and o is not None and all(x != 0 for x in o) and sum(1 for x in d if x == 0) == 1:
            rest = int(np.prod([x for x in d if x != 0])) if len(d) > 1 else 1
            tot = int(np.prod(o))
            if rest > 0 and tot % rest == 0:
                s.put(ins[0], [x if x != 0 else tot // rest for x in d], where)
    elif op in ("slice", "crop", "slice_axis"):
        d = s.get(ins[0])
        if d is not None:
            if op == "slice_axis":
                ax = int(a["axis"]) % len(d)
                begin = [None] * ax + [a.get("begin")]
                end = [None] * ax + [a.get("end")]
                step = []
            else:
                begin, end, step = list(a.get("begin", ())), list(a.get("end", ())), list(a.get("step", ()) or ())
            od = []
            for i, n in enumerate(d):
                if n == 0 or i >= len(begin):
                    od.append(n)
                    continue
                none = lambda v: None if v in (None, "None") else int(v)     # noqa: E731
                st = none(step[i]) if i < len(step) else None
                od.append(len(range(*slice(none(begin[i]), none(end[i]) if i < len(end) else None, st).indices(n))))
            s.put(out0, od, where)
    elif op in ("Embedding", "_contrib_SparseEmbedding"):
        d = s.get(ins[0])
        if d is not None:
            s.put(out0, list(d) + [int(a["output_dim"])], where)
        if len(ins) > 1:
            s.put(ins[1], [int(a["input_dim"]), int(a["output_dim"])], where)


def solve(sym, shapes, partial_heads=False):
    """``shapes``: {variable name: tuple with 0 for unknown dimensions}.  Returns {variable name: fully known shape} for
    every variable whose shape could be completed (including the ones that were complete already), and the list of the
    symbol's output shapes (None where still unknown; with ``partial_heads`` a partly known shape is returned with its
    0 dimensions, as ``infer_shape_partial`` reports it)."""
    order = sym._topo()
    s = _Solver()
    for node in order:
        if node.op is None:
            shp = shapes.get(node.name)
            if shp is None and "__shape__" in node.var_attrs:
                shp = tuple(node.var_attrs["__shape__"])
            if shp is not None:
                s.S[(id(node), 0)] = [int(d) for d in shp]
    for _sweep in range(16):
        s.changed = False
        for node in order:
            if node.op is not None:
                _apply(node, s)
        for node in reversed(order):
            if node.op is not None:
                _apply(node, s)
        if not s.changed:
            break
    out = {}
    for node in order:
        if node.op is None:
            d = s.S.get((id(node), 0))
            if d is not None and all(x != 0 for x in d):
                out[node.name] = tuple(d)
    heads = []
    for n, i in sym._outputs:
        d = s.S.get((id(n), i))
        heads.append(tuple(d) if d is not None and (partial_heads or all(x != 0 for x in d)) else None)
    return out, heads
