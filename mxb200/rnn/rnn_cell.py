"""Symbolic recurrent cells: each call adds the nodes of one time step to a ``mx.sym`` graph; ``unroll`` builds the whole
sequence (reference: python/mxnet/rnn/rnn_cell.py).  Parameters are shared across steps through ``RNNParams``.

Provenance: adapted from Apache MXNet (incubating) 1.3.1 ``python/mxnet/rnn/rnn_cell.py`` (Apache License 2.0).  The public API, argument
names and control flow follow the original so that user scripts keep working; the tensor maths runs on torch /
the mxb200 kernels.  This file is API-compatibility glue, not original B200 work ("Provenance").
"""
from .. import symbol
from ..base import MXNetError

__all__ = ["RNNParams", "BaseRNNCell", "RNNCell", "LSTMCell", "GRUCell", "FusedRNNCell", "SequentialRNNCell", "DropoutCell",
           "ModifierCell", "ZoneoutCell", "ResidualCell", "BidirectionalCell", "BaseConvRNNCell", "ConvRNNCell", "ConvLSTMCell",
           "ConvGRUCell"]


class RNNParams(object):
    """Container of the weight variables of a cell, created lazily by name (rnn_cell.py:78-105)."""

    def __init__(self, prefix=""):
        self._prefix = prefix
        self._params = {}

    def get(self, name, **kwargs):
        name = self._prefix + name
        if name not in self._params:
            self._params[name] = symbol.Variable(name, **kwargs)
        return self._params[name]


def _normalize_sequence(length, inputs, layout, merge, in_layout=None):
    assert inputs is not None, "unroll(inputs=None) has been deprecated. Please create input variables outside unroll."
    axis = layout.find("T")
    in_axis = in_layout.find("T") if in_layout is not None else axis
    if isinstance(inputs, symbol.Symbol):
        if merge is False:
            assert len(inputs.list_outputs()) == 1, "unroll doesn't allow grouped symbol as input."
            inputs = list(symbol.split(inputs, axis=in_axis, num_outputs=length, squeeze_axis=1))
            if length == 1:
                inputs = [inputs] if isinstance(inputs, symbol.Symbol) else inputs
    else:
        assert length is None or len(inputs) == length
        if merge is True:
            inputs = [symbol.expand_dims(i, axis=axis) for i in inputs]
            inputs = symbol.concat(*inputs, dim=axis)
            in_axis = axis
    if isinstance(inputs, symbol.Symbol) and axis != in_axis:
        inputs = symbol.swapaxes(inputs, dim1=axis, dim2=in_axis)
    return inputs, axis


class BaseRNNCell(object):
    def __init__(self, prefix="", params=None):
        if params is None:
            params = RNNParams(prefix)
            self._own_params = True
        else:
            self._own_params = False
        self._prefix, self._params, self._modified = prefix, params, False
        self.reset()

    def reset(self):
        self._init_counter = -1
        self._counter = -1

    def __call__(self, inputs, states):
        raise NotImplementedError()

    @property
    def params(self):
        self._own_params = False
        return self._params

    @property
    def state_info(self):
        raise NotImplementedError()

    @property
    def state_shape(self):
        return [ele["shape"] for ele in self.state_info]

    @property
    def _gate_names(self):
        return ()

    def begin_state(self, func=symbol.zeros, **kwargs):
        assert not self._modified, "After applying modifier cells the base cell cannot be called directly. Call the modifier cell instead."
        states = []
        for info in self.state_info:
            self._init_counter += 1
            kw = dict(kwargs)
            if info is not None:
                kw.update({k: v for k, v in info.items() if k != "__layout__"})
            shape = kw.pop("shape", None)
            state = symbol.var("%sbegin_state_%d" % (self._prefix, self._init_counter), shape=shape) if func is symbol.var \
                else func(shape=tuple(1 if d == 0 else d for d in shape), **kw) if shape is not None else func(**kw)
            states.append(state)
        return states

    def unpack_weights(self, args):
        """fused i2h/h2h matrices -> per-gate matrices (name scheme ``<prefix>i2h<gate>_weight``)."""
        args = args.copy()
        if not self._gate_names:
            return args
        h = self._num_hidden
        for group in ("i2h", "h2h"):
            w = args.pop("%s%s_weight" % (self._prefix, group))
            b = args.pop("%s%s_bias" % (self._prefix, group))
            for j, gate in enumerate(self._gate_names):
                args["%s%s%s_weight" % (self._prefix, group, gate)] = w[j * h:(j + 1) * h].copy()
                args["%s%s%s_bias" % (self._prefix, group, gate)] = b[j * h:(j + 1) * h].copy()
        return args

    def pack_weights(self, args):
        from .. import ndarray as nd
        args = args.copy()
        if not self._gate_names:
            return args
        for group in ("i2h", "h2h"):
            ws, bs = [], []
            for gate in self._gate_names:
                ws.append(args.pop("%s%s%s_weight" % (self._prefix, group, gate)))
                bs.append(args.pop("%s%s%s_bias" % (self._prefix, group, gate)))
            args["%s%s_weight" % (self._prefix, group)] = nd.concat(*ws, dim=0)
            args["%s%s_bias" % (self._prefix, group)] = nd.concat(*bs, dim=0)
        return args

    def _default_states(self, x):
        """Zero begin states whose batch size follows the step input ``x`` [N, ...] (the reference leaves the batch
        dimension 0 = unknown and lets shape inference fill it in; here it is taken from the data)."""
        col = symbol.sum(symbol.reshape(x, shape=(0, -1)), axis=1, keepdims=True) * 0.0          # [N, 1] zeros
        states = []
        for info in self.state_info:
            shape = tuple(info["shape"])
            b = shape.index(0) if 0 in shape else 0                  # the batch axis: 0 for cells, 1 for the fused [L, N, H] state
            unit = tuple(1 if i == b else d for i, d in enumerate(shape))
            n_shape = tuple(-1 if i == b else 1 for i in range(len(shape)))
            states.append(symbol.broadcast_add(symbol.zeros(shape=unit), symbol.reshape(col, shape=n_shape)))
        return states

    def unroll(self, length, inputs, begin_state=None, layout="NTC", merge_outputs=None):
        self.reset()
        inputs, _ = _normalize_sequence(length, inputs, layout, False)
        if begin_state is None:
            begin_state = self._default_states(inputs[0])
        states = begin_state
        outputs = []
        for i in range(length):
            output, states = self(inputs[i], states)
            outputs.append(output)
        outputs, _ = _normalize_sequence(length, outputs, layout, merge_outputs)
        return outputs, states

    def _get_activation(self, inputs, activation, **kwargs):
        if isinstance(activation, str):
            return symbol.Activation(inputs, act_type=activation, **kwargs)
        return activation(inputs, **kwargs)


class RNNCell(BaseRNNCell):
    def __init__(self, num_hidden, activation="tanh", prefix="rnn_", params=None):
        super(RNNCell, self).__init__(prefix=prefix, params=params)
        self._num_hidden, self._activation = num_hidden, activation
        self._iW, self._iB = self.params.get("i2h_weight"), self.params.get("i2h_bias")
        self._hW, self._hB = self.params.get("h2h_weight"), self.params.get("h2h_bias")

    @property
    def state_info(self):
        return [{"shape": (0, self._num_hidden), "__layout__": "NC"}]

    @property
    def _gate_names(self):
        return ("",)

    def __call__(self, inputs, states):
        self._counter += 1
        name = "%st%d_" % (self._prefix, self._counter)
        i2h = symbol.FullyConnected(data=inputs, weight=self._iW, bias=self._iB, num_hidden=self._num_hidden, name="%si2h" % name)
        h2h = symbol.FullyConnected(data=states[0], weight=self._hW, bias=self._hB, num_hidden=self._num_hidden, name="%sh2h" % name)
        output = self._get_activation(i2h + h2h, self._activation, name="%sout" % name)
        return output, [output]


class LSTMCell(BaseRNNCell):
    def __init__(self, num_hidden, prefix="lstm_", params=None, forget_bias=1.0):
        super(LSTMCell, self).__init__(prefix=prefix, params=params)
        self._num_hidden = num_hidden
        from .. import initializer as _init
        self._iW, self._hW = self.params.get("i2h_weight"), self.params.get("h2h_weight")
        # the forget-gate slice of the input bias starts at forget_bias (rnn_cell.py:408-412 of the reference)
        self._iB = self.params.get("i2h_bias", init=_init.LSTMBias(forget_bias=forget_bias))
        self._hB = self.params.get("h2h_bias")
        self._forget_bias = forget_bias

    @property
    def state_info(self):
        return [{"shape": (0, self._num_hidden), "__layout__": "NC"}, {"shape": (0, self._num_hidden), "__layout__": "NC"}]

    @property
    def _gate_names(self):
        return ["_i", "_f", "_c", "_o"]

    def __call__(self, inputs, states):
        self._counter += 1
        name = "%st%d_" % (self._prefix, self._counter)
        i2h = symbol.FullyConnected(data=inputs, weight=self._iW, bias=self._iB, num_hidden=self._num_hidden * 4, name="%si2h" % name)
        h2h = symbol.FullyConnected(data=states[0], weight=self._hW, bias=self._hB, num_hidden=self._num_hidden * 4, name="%sh2h" % name)
        gates = symbol.SliceChannel(i2h + h2h, num_outputs=4, name="%sslice" % name)
        in_gate = symbol.Activation(gates[0], act_type="sigmoid", name="%si" % name)
        forget_gate = symbol.Activation(gates[1], act_type="sigmoid", name="%sf" % name)
        in_transform = symbol.Activation(gates[2], act_type="tanh", name="%sc" % name)
        out_gate = symbol.Activation(gates[3], act_type="sigmoid", name="%so" % name)
        next_c = symbol._internal._plus(forget_gate * states[1], in_gate * in_transform, name="%sstate" % name)
        next_h = symbol._internal._mul(out_gate, symbol.Activation(next_c, act_type="tanh"), name="%sout" % name)
        return next_h, [next_h, next_c]


class GRUCell(BaseRNNCell):
    def __init__(self, num_hidden, prefix="gru_", params=None):
        super(GRUCell, self).__init__(prefix=prefix, params=params)
        self._num_hidden = num_hidden
        self._iW, self._iB = self.params.get("i2h_weight"), self.params.get("i2h_bias")
        self._hW, self._hB = self.params.get("h2h_weight"), self.params.get("h2h_bias")

    @property
    def state_info(self):
        return [{"shape": (0, self._num_hidden), "__layout__": "NC"}]

    @property
    def _gate_names(self):
        return ["_r", "_z", "_o"]

    def __call__(self, inputs, states):
        self._counter += 1
        name = "%st%d_" % (self._prefix, self._counter)
        prev = states[0]
        i2h = symbol.FullyConnected(data=inputs, weight=self._iW, bias=self._iB, num_hidden=self._num_hidden * 3, name="%si2h" % name)
        h2h = symbol.FullyConnected(data=prev, weight=self._hW, bias=self._hB, num_hidden=self._num_hidden * 3, name="%sh2h" % name)
        i2h_r, i2h_z, i2h = symbol.SliceChannel(i2h, num_outputs=3, name="%si2h_slice" % name)
        h2h_r, h2h_z, h2h = symbol.SliceChannel(h2h, num_outputs=3, name="%sh2h_slice" % name)
        reset_gate = symbol.Activation(i2h_r + h2h_r, act_type="sigmoid", name="%sr_act" % name)
        update_gate = symbol.Activation(i2h_z + h2h_z, act_type="sigmoid", name="%sz_act" % name)
        next_h_tmp = symbol.Activation(i2h + reset_gate * h2h, act_type="tanh", name="%sh_act" % name)
        next_h = symbol._internal._plus((1.0 - update_gate) * next_h_tmp, update_gate * prev, name="%sout" % name)
        return next_h, [next_h]


class SequentialRNNCell(BaseRNNCell):
    def __init__(self, params=None):
        super(SequentialRNNCell, self).__init__(prefix="", params=params)
        self._override_cell_params = params is not None
        self._cells = []

    def add(self, cell):
        self._cells.append(cell)
        if self._override_cell_params:
            assert cell._own_params, "Either specify params for SequentialRNNCell or child cells, not both."
            cell.params._params.update(self.params._params)
        self.params._params.update(cell.params._params)

    @property
    def state_info(self):
        return sum([c.state_info for c in self._cells], [])

    def begin_state(self, **kwargs):
        assert not self._modified
        return sum([c.begin_state(**kwargs) for c in self._cells], [])

    def unpack_weights(self, args):
        for c in self._cells:
            args = c.unpack_weights(args)
        return args

    def pack_weights(self, args):
        for c in self._cells:
            args = c.pack_weights(args)
        return args

    def __call__(self, inputs, states):
        self._counter += 1
        next_states = []
        p = 0
        for cell in self._cells:
            assert not isinstance(cell, BidirectionalCell)
            n = len(cell.state_info)
            inputs, st = cell(inputs, states[p:p + n])
            p += n
            next_states.append(st)
        return inputs, sum(next_states, [])

    def unroll(self, length, inputs, begin_state=None, layout="NTC", merge_outputs=None):
        self.reset()
        num_cells = len(self._cells)
        if begin_state is None:
            first, _ = _normalize_sequence(length, inputs, layout, False)
            begin_state = sum([c._default_states(first[0]) for c in self._cells], [])
        p = 0
        next_states = []
        for i, cell in enumerate(self._cells):
            n = len(cell.state_info)
            inputs, states = cell.unroll(length, inputs=inputs, begin_state=begin_state[p:p + n], layout=layout,
                                         merge_outputs=None if i < num_cells - 1 else merge_outputs)
            p += n
            next_states.extend(states)
        return inputs, next_states


class DropoutCell(BaseRNNCell):
    def __init__(self, dropout, prefix="dropout_", params=None):
        super(DropoutCell, self).__init__(prefix, params)
        self.dropout = dropout

    @property
    def state_info(self):
        return []

    def __call__(self, inputs, states):
        if self.dropout > 0:
            inputs = symbol.Dropout(data=inputs, p=self.dropout)
        return inputs, states

    def unroll(self, length, inputs, begin_state=None, layout="NTC", merge_outputs=None):
        self.reset()
        inputs, _ = _normalize_sequence(length, inputs, layout, merge_outputs)
        if isinstance(inputs, symbol.Symbol):
            return self(inputs, [])
        return super(DropoutCell, self).unroll(length, inputs, begin_state=begin_state, layout=layout, merge_outputs=merge_outputs)


class ModifierCell(BaseRNNCell):
    def __init__(self, base_cell):
        super(ModifierCell, self).__init__()
        base_cell._modified = True
        self.base_cell = base_cell

    @property
    def params(self):
        self._own_params = False
        return self.base_cell.params

    @property
    def state_info(self):
        return self.base_cell.state_info

    def begin_state(self, init_sym=symbol.zeros, **kwargs):
        assert not self._modified
        self.base_cell._modified = False
        begin = self.base_cell.begin_state(func=init_sym, **kwargs)
        self.base_cell._modified = True
        return begin

    def unpack_weights(self, args):
        return self.base_cell.unpack_weights(args)

    def pack_weights(self, args):
        return self.base_cell.pack_weights(args)


class ZoneoutCell(ModifierCell):
    def __init__(self, base_cell, zoneout_outputs=0., zoneout_states=0.):
        assert not isinstance(base_cell, FusedRNNCell), "FusedRNNCell doesn't support zoneout. Please unfuse first."
        assert not isinstance(base_cell, BidirectionalCell), "BidirectionalCell doesn't support zoneout since it doesn't support step."
        super(ZoneoutCell, self).__init__(base_cell)
        self.zoneout_outputs, self.zoneout_states = zoneout_outputs, zoneout_states
        self.prev_output = None

    def reset(self):
        super(ZoneoutCell, self).reset()
        self.prev_output = None

    def __call__(self, inputs, states):
        cell, p_out, p_st = self.base_cell, self.zoneout_outputs, self.zoneout_states
        next_output, next_states = cell(inputs, states)

        def mask(p, like):
            return symbol.Dropout(symbol.ones_like(like), p=p)
        prev_output = self.prev_output if self.prev_output is not None else symbol.zeros_like(next_output)
        output = symbol.where(mask(p_out, next_output), next_output, prev_output) if p_out != 0. else next_output
        new_states = [symbol.where(mask(p_st, ns), ns, os) for ns, os in zip(next_states, states)] if p_st != 0. else next_states
        self.prev_output = output
        return output, new_states


class ResidualCell(ModifierCell):
    def __call__(self, inputs, states):
        output, states = self.base_cell(inputs, states)
        return symbol.elemwise_add(output, inputs, name="%s_plus_residual" % output.name), states

    def unroll(self, length, inputs, begin_state=None, layout="NTC", merge_outputs=None):
        self.reset()
        self.base_cell._modified = False
        outputs, states = self.base_cell.unroll(length, inputs=inputs, begin_state=begin_state, layout=layout, merge_outputs=merge_outputs)
        self.base_cell._modified = True
        merge_outputs = isinstance(outputs, symbol.Symbol) if merge_outputs is None else merge_outputs
        inputs, _ = _normalize_sequence(length, inputs, layout, merge_outputs)
        if merge_outputs:
            outputs = symbol.elemwise_add(outputs, inputs)
        else:
            outputs = [symbol.elemwise_add(o, i) for o, i in zip(outputs, inputs)]
        return outputs, states


class BidirectionalCell(BaseRNNCell):
    def __init__(self, l_cell, r_cell, params=None, output_prefix="bi_"):
        super(BidirectionalCell, self).__init__("", params=params)
        self._output_prefix = output_prefix
        self._override_cell_params = params is not None
        if self._override_cell_params:
            assert l_cell._own_params and r_cell._own_params
            l_cell.params._params.update(self.params._params)
            r_cell.params._params.update(self.params._params)
        self.params._params.update(l_cell.params._params)
        self.params._params.update(r_cell.params._params)
        self._cells = [l_cell, r_cell]

    def unpack_weights(self, args):
        for c in self._cells:
            args = c.unpack_weights(args)
        return args

    def pack_weights(self, args):
        for c in self._cells:
            args = c.pack_weights(args)
        return args

    def __call__(self, inputs, states):
        raise NotImplementedError("Bidirectional cannot be stepped. Please use unroll")

    @property
    def state_info(self):
        return sum([c.state_info for c in self._cells], [])

    def begin_state(self, **kwargs):
        assert not self._modified
        return sum([c.begin_state(**kwargs) for c in self._cells], [])

    def unroll(self, length, inputs, begin_state=None, layout="NTC", merge_outputs=None):
        self.reset()
        inputs, axis = _normalize_sequence(length, inputs, layout, False)
        if begin_state is None:
            begin_state = sum([c._default_states(inputs[0]) for c in self._cells], [])
        l_cell, r_cell = self._cells
        nl = len(l_cell.state_info)
        l_out, l_st = l_cell.unroll(length, inputs=inputs, begin_state=begin_state[:nl], layout=layout, merge_outputs=False)
        r_out, r_st = r_cell.unroll(length, inputs=list(reversed(inputs)), begin_state=begin_state[nl:], layout=layout, merge_outputs=False)
        outputs = [symbol.concat(lo, ro, dim=1, name="%st%d" % (self._output_prefix, i))
                   for i, (lo, ro) in enumerate(zip(l_out, reversed(r_out)))]
        if merge_outputs:
            outputs, _ = _normalize_sequence(length, outputs, layout, True)
        return outputs, l_st + r_st


class FusedRNNCell(BaseRNNCell):
    """Multi-layer (bi)directional RNN/LSTM/GRU as ONE ``RNN`` operator over the whole sequence with a single packed
    parameter vector (cuDNN layout; rnn_cell.py:536-745).  ``unfuse()`` returns the equivalent stack of step cells."""

    def __init__(self, num_hidden, num_layers=1, mode="lstm", bidirectional=False, dropout=0., get_next_state=False,
                 forget_bias=1.0, prefix=None, params=None):
        if prefix is None:
            prefix = "%s_" % mode
        super(FusedRNNCell, self).__init__(prefix=prefix, params=params)
        self._num_hidden, self._num_layers, self._mode = num_hidden, num_layers, mode
        self._bidirectional, self._dropout, self._get_next_state = bidirectional, dropout, get_next_state
        self._directions = ["l", "r"] if bidirectional else ["l"]
        from .. import initializer as _init
        self._parameter = self.params.get("parameters", init=_init.FusedRNN(None, num_hidden, num_layers, mode, bidirectional,
                                                                            forget_bias))

    @property
    def state_info(self):
        b = self._bidirectional + 1
        n = (self._mode == "lstm") + 1
        return [{"shape": (b * self._num_layers, 0, self._num_hidden), "__layout__": "LNC"} for _ in range(n)]

    @property
    def _gate_names(self):
        return {"rnn_relu": [""], "rnn_tanh": [""], "lstm": ["_i", "_f", "_c", "_o"], "gru": ["_r", "_z", "_o"]}[self._mode]

    @property
    def _num_gates(self):
        return len(self._gate_names)

    def _slice_weights(self, arr, li, lh):
        """name -> view of the flat cuDNN-layout vector ``arr`` (1-D NDArray): all weights layer-major (i2h then h2h, gates
        stacked), then all biases in the same order (rnn_cell.py:600-640 of the reference)."""
        args = {}
        g = self._gate_names
        ng, H = len(g), self._num_hidden
        b = len(self._directions)
        p = 0
        flat = arr
        for layer in range(self._num_layers):
            for d in self._directions:
                inp = li if layer == 0 else b * lh
                for grp, width in (("i2h", inp), ("h2h", lh)):
                    for gate in g:
                        name = "%s%s%d_%s%s_weight" % (self._prefix, d, layer, grp, gate)
                        args[name] = flat[p:p + H * width].reshape((H, width))
                        p += H * width
        for layer in range(self._num_layers):
            for d in self._directions:
                for grp in ("i2h", "h2h"):
                    for gate in g:
                        name = "%s%s%d_%s%s_bias" % (self._prefix, d, layer, grp, gate)
                        args[name] = flat[p:p + H]
                        p += H
        assert p == arr.size, "Invalid parameters size for FusedRNNCell"
        return args

    def unpack_weights(self, args):
        args = args.copy()
        arr = args.pop(self._parameter.name)
        b = len(self._directions)
        m = self._num_gates
        h = self._num_hidden
        num_input = arr.size // b // h // m - (self._num_layers - 1) * (h + b * h + 2) - h - 2
        for name, w in self._slice_weights(arr, num_input, h).items():
            args[name] = w.copy()
        return args

    def pack_weights(self, args):
        from .. import ndarray as nd
        args = args.copy()
        b = len(self._directions)
        m = self._num_gates
        h = self._num_hidden
        w0 = args["%sl0_i2h%s_weight" % (self._prefix, self._gate_names[0])]
        num_input = w0.shape[1]
        total = (num_input + h + 2) * h * m * b + (self._num_layers - 1) * m * h * (h + b * h + 2) * b
        arr = nd.zeros((total,), ctx=w0.context, dtype=w0.dtype)
        for name, view in self._slice_weights(arr, num_input, h).items():
            view[:] = args.pop(name)
        args[self._parameter.name] = arr
        return args

    def __call__(self, inputs, states):
        raise NotImplementedError("FusedRNNCell cannot be stepped. Please use unroll")

    def unroll(self, length, inputs, begin_state=None, layout="NTC", merge_outputs=None):
        self.reset()
        inputs, axis = _normalize_sequence(length, inputs, layout, True)
        if axis == 1:                       # the RNN op is time-major
            inputs = symbol.swapaxes(inputs, dim1=0, dim2=1)
        if begin_state is None:
            begin_state = self._default_states(symbol.squeeze(symbol.slice_axis(inputs, axis=0, begin=0, end=1), axis=0))
        states = {"state": begin_state[0]}
        if self._mode == "lstm":
            states["state_cell"] = begin_state[1]
        rnn = symbol.RNN(data=inputs, parameters=self._parameter, state_size=self._num_hidden, num_layers=self._num_layers,
                         bidirectional=self._bidirectional, p=self._dropout, state_outputs=self._get_next_state,
                         mode=self._mode, name=self._prefix + "rnn", **states)
        if not self._get_next_state:
            outputs, states = rnn, []
        elif self._mode == "lstm":
            outputs, states = rnn[0], [rnn[1], rnn[2]]
        else:
            outputs, states = rnn[0], [rnn[1]]
        if axis == 1:
            outputs = symbol.swapaxes(outputs, dim1=0, dim2=1)
        outputs, _ = _normalize_sequence(length, outputs, layout, merge_outputs)
        return outputs, states

    def unfuse(self):
        stack = SequentialRNNCell()
        get_cell = {"rnn_relu": lambda p: RNNCell(self._num_hidden, activation="relu", prefix=p),
                    "rnn_tanh": lambda p: RNNCell(self._num_hidden, activation="tanh", prefix=p),
                    "lstm": lambda p: LSTMCell(self._num_hidden, prefix=p),
                    "gru": lambda p: GRUCell(self._num_hidden, prefix=p)}[self._mode]
        for i in range(self._num_layers):
            if self._bidirectional:
                stack.add(BidirectionalCell(get_cell("%sl%d_" % (self._prefix, i)), get_cell("%sr%d_" % (self._prefix, i)),
                                            output_prefix="%sbi_l%d_" % (self._prefix, i)))
            else:
                stack.add(get_cell("%sl%d_" % (self._prefix, i)))
            if self._dropout > 0 and i != self._num_layers - 1:
                stack.add(DropoutCell(self._dropout, prefix="%s_dropout%d_" % (self._prefix, i)))
        return stack


# ---------------------------------------------------------------------------------------------------- convolutional cells
class BaseConvRNNCell(BaseRNNCell):
    """Cells whose input-to-hidden and hidden-to-hidden transforms are convolutions (Shi et al. 2015; reference:
    python/mxnet/rnn/rnn_cell.py:1094-1190).  ``input_shape`` is the per-step input (N, C, spatial...); the state has the
    shape of the i2h convolution's output with ``num_hidden`` channels; the h2h convolution is padded to keep that shape."""

    def __init__(self, input_shape, num_hidden, h2h_kernel, h2h_dilate, i2h_kernel, i2h_stride, i2h_pad, i2h_dilate,
                 i2h_weight_initializer, h2h_weight_initializer, i2h_bias_initializer, h2h_bias_initializer, activation, prefix="",
                 params=None, conv_layout="NCHW"):
        super(BaseConvRNNCell, self).__init__(prefix=prefix, params=params)
        self._h2h_kernel, self._h2h_dilate = tuple(h2h_kernel), tuple(h2h_dilate)
        assert all(k % 2 == 1 for k in self._h2h_kernel), "Only support odd number, get h2h_kernel= %s" % str(h2h_kernel)
        self._h2h_pad = tuple(d * (k - 1) // 2 for k, d in zip(self._h2h_kernel, self._h2h_dilate))
        self._i2h_kernel, self._i2h_stride = tuple(i2h_kernel), tuple(i2h_stride)
        self._i2h_pad, self._i2h_dilate = tuple(i2h_pad), tuple(i2h_dilate)
        self._num_hidden, self._input_shape, self._conv_layout = num_hidden, tuple(input_shape), conv_layout
        self._activation = activation
        # state shape = output shape of the i2h convolution
        data = symbol.Variable("data")
        conv = symbol.Convolution(data=data, num_filter=num_hidden, kernel=self._i2h_kernel, stride=self._i2h_stride,
                                  pad=self._i2h_pad, dilate=self._i2h_dilate, layout=conv_layout)
        self._state_shape = (0,) + tuple(conv.infer_shape(data=self._input_shape)[1][0][1:])
        self._iW = self.params.get("i2h_weight", init=i2h_weight_initializer)
        self._hW = self.params.get("h2h_weight", init=h2h_weight_initializer)
        self._iB = self.params.get("i2h_bias", init=i2h_bias_initializer)
        self._hB = self.params.get("h2h_bias", init=h2h_bias_initializer)

    @property
    def _num_gates(self):
        return len(self._gate_names)

    @property
    def state_info(self):
        return [{"shape": self._state_shape, "__layout__": self._conv_layout},
                {"shape": self._state_shape, "__layout__": self._conv_layout}]

    def _conv_forward(self, inputs, states, name):
        nf = self._num_hidden * self._num_gates
        i2h = symbol.Convolution(name="%si2h" % name, data=inputs, num_filter=nf, kernel=self._i2h_kernel, stride=self._i2h_stride,
                                 pad=self._i2h_pad, dilate=self._i2h_dilate, weight=self._iW, bias=self._iB, layout=self._conv_layout)
        h2h = symbol.Convolution(name="%sh2h" % name, data=states[0], num_filter=nf, kernel=self._h2h_kernel, dilate=self._h2h_dilate,
                                 pad=self._h2h_pad, stride=(1,) * len(self._h2h_kernel), weight=self._hW, bias=self._hB,
                                 layout=self._conv_layout)
        return i2h, h2h

    def __call__(self, inputs, states):
        raise NotImplementedError("BaseConvRNNCell is abstract class for convolutional RNN")


class ConvRNNCell(BaseConvRNNCell):
    def __init__(self, input_shape, num_hidden, h2h_kernel=(3, 3), h2h_dilate=(1, 1), i2h_kernel=(3, 3), i2h_stride=(1, 1),
                 i2h_pad=(1, 1), i2h_dilate=(1, 1), i2h_weight_initializer=None, h2h_weight_initializer=None,
                 i2h_bias_initializer="zeros", h2h_bias_initializer="zeros", activation=None, prefix="ConvRNN_", params=None,
                 conv_layout="NCHW"):
        act = activation if activation is not None else (lambda x, **kw: symbol.LeakyReLU(x, act_type="leaky", slope=0.2, **kw))
        super(ConvRNNCell, self).__init__(input_shape, num_hidden, h2h_kernel, h2h_dilate, i2h_kernel, i2h_stride, i2h_pad, i2h_dilate,
                                          i2h_weight_initializer, h2h_weight_initializer, i2h_bias_initializer, h2h_bias_initializer,
                                          act, prefix=prefix, params=params, conv_layout=conv_layout)

    @property
    def _gate_names(self):
        return ("",)

    @property
    def state_info(self):
        return [{"shape": self._state_shape, "__layout__": self._conv_layout}]

    def __call__(self, inputs, states):
        self._counter += 1
        name = "%st%d_" % (self._prefix, self._counter)
        i2h, h2h = self._conv_forward(inputs, states, name)
        output = self._get_activation(i2h + h2h, self._activation, name="%sout" % name)
        return output, [output]


class ConvLSTMCell(BaseConvRNNCell):
    def __init__(self, input_shape, num_hidden, h2h_kernel=(3, 3), h2h_dilate=(1, 1), i2h_kernel=(3, 3), i2h_stride=(1, 1),
                 i2h_pad=(1, 1), i2h_dilate=(1, 1), i2h_weight_initializer=None, h2h_weight_initializer=None,
                 i2h_bias_initializer="zeros", h2h_bias_initializer="zeros", activation=None, prefix="ConvLSTM_", params=None,
                 conv_layout="NCHW"):
        act = activation if activation is not None else (lambda x, **kw: symbol.LeakyReLU(x, act_type="leaky", slope=0.2, **kw))
        super(ConvLSTMCell, self).__init__(input_shape, num_hidden, h2h_kernel, h2h_dilate, i2h_kernel, i2h_stride, i2h_pad, i2h_dilate,
                                           i2h_weight_initializer, h2h_weight_initializer, i2h_bias_initializer, h2h_bias_initializer,
                                           act, prefix=prefix, params=params, conv_layout=conv_layout)

    @property
    def _gate_names(self):
        return ["_i", "_f", "_c", "_o"]

    def __call__(self, inputs, states):
        self._counter += 1
        name = "%st%d_" % (self._prefix, self._counter)
        i2h, h2h = self._conv_forward(inputs, states, name)
        ch = self._conv_layout.find("C")
        gates = symbol.SliceChannel(i2h + h2h, num_outputs=4, axis=ch, name="%sslice" % name)
        in_gate = symbol.Activation(gates[0], act_type="sigmoid", name="%si" % name)
        forget_gate = symbol.Activation(gates[1], act_type="sigmoid", name="%sf" % name)
        in_transform = self._get_activation(gates[2], self._activation, name="%sc" % name)
        out_gate = symbol.Activation(gates[3], act_type="sigmoid", name="%so" % name)
        next_c = symbol._internal._plus(forget_gate * states[1], in_gate * in_transform, name="%sstate" % name)
        next_h = symbol._internal._mul(out_gate, self._get_activation(next_c, self._activation), name="%sout" % name)
        return next_h, [next_h, next_c]


class ConvGRUCell(BaseConvRNNCell):
    def __init__(self, input_shape, num_hidden, h2h_kernel=(3, 3), h2h_dilate=(1, 1), i2h_kernel=(3, 3), i2h_stride=(1, 1),
                 i2h_pad=(1, 1), i2h_dilate=(1, 1), i2h_weight_initializer=None, h2h_weight_initializer=None,
                 i2h_bias_initializer="zeros", h2h_bias_initializer="zeros", activation=None, prefix="ConvGRU_", params=None,
                 conv_layout="NCHW"):
        act = activation if activation is not None else (lambda x, **kw: symbol.LeakyReLU(x, act_type="leaky", slope=0.2, **kw))
        super(ConvGRUCell, self).__init__(input_shape, num_hidden, h2h_kernel, h2h_dilate, i2h_kernel, i2h_stride, i2h_pad, i2h_dilate,
                                          i2h_weight_initializer, h2h_weight_initializer, i2h_bias_initializer, h2h_bias_initializer,
                                          act, prefix=prefix, params=params, conv_layout=conv_layout)

    @property
    def _gate_names(self):
        return ["_r", "_z", "_o"]

    @property
    def state_info(self):
        return [{"shape": self._state_shape, "__layout__": self._conv_layout}]

    def __call__(self, inputs, states):
        self._counter += 1
        name = "%st%d_" % (self._prefix, self._counter)
        i2h, h2h = self._conv_forward(inputs, states, name)
        ch = self._conv_layout.find("C")
        i2h_r, i2h_z, i2h = symbol.SliceChannel(i2h, num_outputs=3, axis=ch, name="%si2h_slice" % name)
        h2h_r, h2h_z, h2h = symbol.SliceChannel(h2h, num_outputs=3, axis=ch, name="%sh2h_slice" % name)
        reset_gate = symbol.Activation(i2h_r + h2h_r, act_type="sigmoid", name="%sr_act" % name)
        update_gate = symbol.Activation(i2h_z + h2h_z, act_type="sigmoid", name="%sz_act" % name)
        next_h_tmp = self._get_activation(i2h + reset_gate * h2h, self._activation, name="%sh_act" % name)
        next_h = symbol._internal._plus((1.0 - update_gate) * next_h_tmp, update_gate * states[0], name="%sout" % name)
        return next_h, [next_h]
