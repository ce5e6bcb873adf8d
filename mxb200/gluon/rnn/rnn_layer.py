"""Fused RNN / LSTM / GRU layers (reference: python/mxnet/gluon/rnn/rnn_layer.py -> src/operator/rnn.cu / cuDNN RNN).
Parameters are kept per layer/direction with the reference names ({l,r}{i}_{i2h,h2h}_{weight,bias}); the
forward pass uses torch's fused (cuDNN) kernels."""
import torch

from ..block import HybridBlock
from ...ndarray import NDArray, zeros
from ...ndarray.ndarray import _op
from ... import autograd

__all__ = ["RNN", "LSTM", "GRU"]

_GATES = {"rnn_relu": 1, "rnn_tanh": 1, "lstm": 4, "gru": 3}


def _run(mode, x, weights, h0, c0, num_layers, bidirectional, dropout, training):
    flat = [w for w in weights]
    if mode == "lstm":
        out, h, c = torch._VF.lstm(x, (h0, c0), flat, True, num_layers, dropout, training, bidirectional, False)
        return out, h, c
    if mode == "gru":
        out, h = torch._VF.gru(x, h0, flat, True, num_layers, dropout, training, bidirectional, False)
    elif mode == "rnn_relu":
        out, h = torch._VF.rnn_relu(x, h0, flat, True, num_layers, dropout, training, bidirectional, False)
    else:
        out, h = torch._VF.rnn_tanh(x, h0, flat, True, num_layers, dropout, training, bidirectional, False)
    return out, h, None


def _fused_rnn_forward(data, parameters, state, state_cell, state_size, num_layers, bidirectional, mode, p, state_outputs):
    """nd.RNN with a flat cuDNN-layout parameter vector: all weights (layer-major, i2h then h2h) then all biases."""
    ndir = 2 if bidirectional else 1
    g = _GATES[mode]
    H = state_size
    I = data.shape[2]
    flat = parameters._t
    shapes_w, shapes_b = [], []
    for l in range(num_layers):
        inp = I if l == 0 else H * ndir
        for _d in range(ndir):
            shapes_w += [(g * H, inp), (g * H, H)]
            shapes_b += [(g * H,), (g * H,)]
    off = 0
    ws, bs = [], []
    for s in shapes_w:
        n = s[0] * s[1]
        ws.append(flat[off:off + n].view(s)); off += n
    for s in shapes_b:
        bs.append(flat[off:off + s[0]]); off += s[0]
    weights = []
    for i in range(0, len(ws), 2):
        weights += [ws[i], ws[i + 1], bs[i], bs[i + 1]]
    out, h, c = _run(mode, data._t, weights, state._t, state_cell._t if state_cell is not None else None, num_layers,
                     bidirectional, p, autograd.is_training())
    outs = [NDArray(out, data.context)]
    if state_outputs:
        outs.append(NDArray(h, data.context))
        if c is not None:
            outs.append(NDArray(c, data.context))
        return outs
    return outs[0]


class _RNNLayer(HybridBlock):
    def __init__(self, hidden_size, num_layers, layout, dropout, bidirectional, input_size, i2h_weight_initializer,
                 h2h_weight_initializer, i2h_bias_initializer, h2h_bias_initializer, mode, **kwargs):
        super(_RNNLayer, self).__init__(**kwargs)
        assert layout in ("TNC", "NTC"), "Invalid layout %s; must be one of ['TNC' or 'NTC']" % layout
        self._hidden_size, self._num_layers, self._mode, self._layout = hidden_size, num_layers, mode, layout
        self._dropout, self._dir, self._input_size = dropout, 2 if bidirectional else 1, input_size
        self._gates = _GATES[mode]
        ng, ni, nh = self._gates, input_size, hidden_size
        self._param_names = []
        for i in range(num_layers):
            for j in ["l", "r"][:self._dir]:
                for nm, shape, init in [("i2h_weight", (ng * nh, ni), i2h_weight_initializer),
                                        ("h2h_weight", (ng * nh, nh), h2h_weight_initializer),
                                        ("i2h_bias", (ng * nh,), i2h_bias_initializer),
                                        ("h2h_bias", (ng * nh,), h2h_bias_initializer)]:
                    name = "%s%d_%s" % (j, i, nm)
                    p = self.params.get(name, shape=shape, init=init, allow_deferred_init=True)
                    self._reg_params[name] = p
                    self._param_names.append(name)
            ni = nh * self._dir

    def __getattr__(self, name):
        # layer.l0_i2h_weight etc.: the per-layer parameters are attributes in the reference
        reg = self.__dict__.get("_reg_params")
        if reg is not None and name in reg:
            return reg[name]
        raise AttributeError("'%s' object has no attribute '%s'" % (type(self).__name__, name))

    def _unfuse(self):
        """The same network as a stack of cells sharing this layer's parameters (rnn_layer.py:115-149 of the reference):
        one (Bidirectional) cell per layer, DropoutCell in between."""
        from . import rnn_cell
        make = {"rnn_relu": lambda **kw: rnn_cell.RNNCell(self._hidden_size, activation="relu", **kw),
                "rnn_tanh": lambda **kw: rnn_cell.RNNCell(self._hidden_size, activation="tanh", **kw),
                "lstm": lambda **kw: rnn_cell.LSTMCell(self._hidden_size, **kw),
                "gru": lambda **kw: rnn_cell.GRUCell(self._hidden_size, **kw)}[self._mode]
        stack = rnn_cell.HybridSequentialRNNCell(prefix=self.prefix, params=self.params)
        with stack.name_scope():
            ni = self._input_size
            for i in range(self._num_layers):
                kw = {"input_size": ni}
                if self._dir == 2:
                    stack.add(rnn_cell.BidirectionalCell(make(prefix="l%d_" % i, **kw), make(prefix="r%d_" % i, **kw)))
                else:
                    stack.add(make(prefix="l%d_" % i, **kw))
                if self._dropout > 0 and i != self._num_layers - 1:
                    stack.add(rnn_cell.DropoutCell(self._dropout))
                ni = self._hidden_size * self._dir
        return stack

    def _infer_param_shapes(self, x, *args):
        ni = x.shape[2]
        ng, nh = self._gates, self._hidden_size
        for i in range(self._num_layers):
            for j in ["l", "r"][:self._dir]:
                self._reg_params["%s%d_i2h_weight" % (j, i)].shape = (ng * nh, ni)
            ni = nh * self._dir

    def state_info(self, batch_size=0):
        raise NotImplementedError

    def begin_state(self, batch_size=0, func=zeros, **kwargs):
        return [func(info["shape"], **kwargs) for info in self.state_info(batch_size)]

    def forward(self, inputs, states=None):
        batch_size = inputs.shape[self._layout.find("N")]
        skip_states = states is None
        if skip_states:
            states = self.begin_state(batch_size, ctx=inputs.context, dtype=inputs.dtype)
        if isinstance(states, NDArray):
            states = [states]
        ctx = inputs.context
        try:
            params = [self._reg_params[n].data(ctx) for n in self._param_names]
        except Exception:  # deferred init
            self._deferred_infer(inputs if self._layout == "TNC" else inputs.swapaxes(0, 1))
            params = [self._reg_params[n].data(ctx) for n in self._param_names]
        training = autograd.is_training()
        mode, nl, bi, dr = self._mode, self._num_layers, self._dir == 2, self._dropout
        ntc = self._layout == "NTC"

        def run(x, *rest):
            nstate = 2 if mode == "lstm" else 1
            st, ws = rest[:nstate], rest[nstate:]
            if ntc:
                x = x.transpose(0, 1)
            out, h, c = _run(mode, x, list(ws), st[0], st[1] if mode == "lstm" else None, nl, bi, dr, training)
            if ntc:
                out = out.transpose(0, 1)
            return [out, h] + ([c] if c is not None else [])
        res = _op(run, inputs, *(list(states) + params))
        out, new_states = res[0], res[1:]
        return out if skip_states else (out, new_states)

    def __call__(self, inputs, states=None):
        return self.forward(inputs, states)


class RNN(_RNNLayer):
    def __init__(self, hidden_size, num_layers=1, activation="relu", layout="TNC", dropout=0, bidirectional=False,
                 i2h_weight_initializer=None, h2h_weight_initializer=None, i2h_bias_initializer="zeros",
                 h2h_bias_initializer="zeros", input_size=0, **kwargs):
        super(RNN, self).__init__(hidden_size, num_layers, layout, dropout, bidirectional, input_size,
                                  i2h_weight_initializer, h2h_weight_initializer, i2h_bias_initializer,
                                  h2h_bias_initializer, "rnn_" + activation, **kwargs)

    def state_info(self, batch_size=0):
        return [{"shape": (self._num_layers * self._dir, batch_size, self._hidden_size), "__layout__": "LNC"}]


class LSTM(_RNNLayer):
    def __init__(self, hidden_size, num_layers=1, layout="TNC", dropout=0, bidirectional=False, input_size=0,
                 i2h_weight_initializer=None, h2h_weight_initializer=None, i2h_bias_initializer="zeros",
                 h2h_bias_initializer="zeros", **kwargs):
        super(LSTM, self).__init__(hidden_size, num_layers, layout, dropout, bidirectional, input_size,
                                   i2h_weight_initializer, h2h_weight_initializer, i2h_bias_initializer,
                                   h2h_bias_initializer, "lstm", **kwargs)

    def state_info(self, batch_size=0):
        return [{"shape": (self._num_layers * self._dir, batch_size, self._hidden_size), "__layout__": "LNC"},
                {"shape": (self._num_layers * self._dir, batch_size, self._hidden_size), "__layout__": "LNC"}]


class GRU(_RNNLayer):
    def __init__(self, hidden_size, num_layers=1, layout="TNC", dropout=0, bidirectional=False, input_size=0,
                 i2h_weight_initializer=None, h2h_weight_initializer=None, i2h_bias_initializer="zeros",
                 h2h_bias_initializer="zeros", **kwargs):
        super(GRU, self).__init__(hidden_size, num_layers, layout, dropout, bidirectional, input_size,
                                  i2h_weight_initializer, h2h_weight_initializer, i2h_bias_initializer,
                                  h2h_bias_initializer, "gru", **kwargs)

    def state_info(self, batch_size=0):
        return [{"shape": (self._num_layers * self._dir, batch_size, self._hidden_size), "__layout__": "LNC"}]
