"""gluon.contrib.rnn (reference: python/mxnet/gluon/contrib/rnn/{rnn_cell,conv_rnn_cell}.py)."""
from ...rnn.rnn_cell import HybridRecurrentCell, RecurrentCell
from .... import ndarray as nd

__all__ = ["VariationalDropoutCell", "LSTMPCell", "Conv1DRNNCell", "Conv2DRNNCell", "Conv3DRNNCell", "Conv1DLSTMCell",
           "Conv2DLSTMCell", "Conv3DLSTMCell", "Conv1DGRUCell", "Conv2DGRUCell", "Conv3DGRUCell"]


class VariationalDropoutCell(RecurrentCell):
    """Applies the SAME dropout mask at every time step to inputs / states / outputs of ``base_cell`` (Gal &
    Ghahramani; reference rnn_cell.py:26-195).  Masks are drawn at the first step after ``reset()``."""

    def __init__(self, base_cell, drop_inputs=0., drop_states=0., drop_outputs=0.):
        super(VariationalDropoutCell, self).__init__(prefix=base_cell.prefix + "vardrop_", params=None)
        self.base_cell = base_cell
        self.drop_inputs, self.drop_states, self.drop_outputs = drop_inputs, drop_states, drop_outputs
        self._masks = {}

    def state_info(self, batch_size=0):
        return self.base_cell.state_info(batch_size)

    def begin_state(self, batch_size=0, **kwargs):
        return self.base_cell.begin_state(batch_size, **kwargs)

    def reset(self):
        super(VariationalDropoutCell, self).reset()
        self._masks = {}

    def _mask(self, key, like, p):
        if p <= 0:
            return None
        if key not in self._masks:
            self._masks[key] = nd.Dropout(nd.ones_like(like), p=p)
        return self._masks[key]

    def __call__(self, inputs, states):
        from .... import autograd
        if autograd.is_training():
            m = self._mask("in", inputs, self.drop_inputs)
            if m is not None:
                inputs = inputs * m
            m = self._mask("st", states[0], self.drop_states)
            if m is not None:
                states = [states[0] * m] + list(states[1:])
        out, states = self.base_cell(inputs, states)
        if autograd.is_training():
            m = self._mask("out", out, self.drop_outputs)
            if m is not None:
                out = out * m
        return out, states


class LSTMPCell(HybridRecurrentCell):
    """LSTM with a recurrent projection layer (Sak et al. 2014; reference rnn_cell.py:197-330): the hidden state fed
    back and emitted is r_t = W_hr h_t with ``projection_size`` units."""

    def __init__(self, hidden_size, projection_size, i2h_weight_initializer=None, h2h_weight_initializer=None,
                 h2r_weight_initializer=None, i2h_bias_initializer="zeros", h2h_bias_initializer="zeros", input_size=0,
                 prefix=None, params=None):
        super(LSTMPCell, self).__init__(prefix=prefix, params=params)
        self._hidden_size, self._input_size, self._projection_size = hidden_size, input_size, projection_size
        self.i2h_weight = self.params.get("i2h_weight", shape=(4 * hidden_size, input_size), init=i2h_weight_initializer, allow_deferred_init=True)
        self.h2h_weight = self.params.get("h2h_weight", shape=(4 * hidden_size, projection_size), init=h2h_weight_initializer, allow_deferred_init=True)
        self.h2r_weight = self.params.get("h2r_weight", shape=(projection_size, hidden_size), init=h2r_weight_initializer, allow_deferred_init=True)
        self.i2h_bias = self.params.get("i2h_bias", shape=(4 * hidden_size,), init=i2h_bias_initializer, allow_deferred_init=True)
        self.h2h_bias = self.params.get("h2h_bias", shape=(4 * hidden_size,), init=h2h_bias_initializer, allow_deferred_init=True)

    def _infer_param_shapes(self, x, *args):
        self.i2h_weight.shape = (4 * self._hidden_size, x.shape[-1])

    def state_info(self, batch_size=0):
        return [{"shape": (batch_size, self._projection_size), "__layout__": "NC"},
                {"shape": (batch_size, self._hidden_size), "__layout__": "NC"}]

    def _alias(self):
        return "lstmp"

    def hybrid_forward(self, F, inputs, states, i2h_weight, h2h_weight, h2r_weight, i2h_bias, h2h_bias):
        prefix = "t%d_" % self._counter            # node names of the reference cell
        i2h = F.FullyConnected(inputs, i2h_weight, i2h_bias, num_hidden=self._hidden_size * 4, name=prefix + "i2h")
        h2h = F.FullyConnected(states[0], h2h_weight, h2h_bias, num_hidden=self._hidden_size * 4, name=prefix + "h2h")
        s = F.split(i2h + h2h, num_outputs=4, axis=1, name=prefix + "slice")
        in_gate, forget_gate = F.sigmoid(s[0]), F.sigmoid(s[1])
        in_transform, out_gate = F.tanh(s[2]), F.sigmoid(s[3])
        next_c = F._internal._plus(forget_gate * states[1], in_gate * in_transform, name=prefix + "state")
        hidden = F._internal._mul(out_gate, F.tanh(next_c), name=prefix + "hidden")
        next_r = F.FullyConnected(hidden, h2r_weight, no_bias=True, num_hidden=self._projection_size, name=prefix + "out")
        return next_r, [next_r, next_c]


def _tup(v, n):
    return (v,) * n if isinstance(v, int) else tuple(v)


class _BaseConvRNNCell(HybridRecurrentCell):
    """Convolutional recurrence: both the input-to-hidden and hidden-to-hidden maps are convolutions (Shi et al.
    2015; reference conv_rnn_cell.py:37-175).  Layout NC(D)(H)W; the h2h convolution is 'same'-padded (odd kernels)."""
    _ngates = 1

    def __init__(self, input_shape, hidden_channels, i2h_kernel, h2h_kernel, i2h_pad=0, i2h_dilate=1, h2h_dilate=1,
                 i2h_weight_initializer=None, h2h_weight_initializer=None, i2h_bias_initializer="zeros",
                 h2h_bias_initializer="zeros", dims=2, conv_layout="NCHW", activation="tanh", prefix=None, params=None):
        super(_BaseConvRNNCell, self).__init__(prefix=prefix, params=params)
        self._dims, self._hidden_channels, self._input_shape = dims, hidden_channels, tuple(input_shape)
        self._activation = activation
        self._i2h_kernel, self._h2h_kernel = _tup(i2h_kernel, dims), _tup(h2h_kernel, dims)
        assert all(k % 2 == 1 for k in self._h2h_kernel), "Only support odd number, get h2h_kernel= %s" % str(h2h_kernel)
        self._i2h_pad, self._i2h_dilate, self._h2h_dilate = _tup(i2h_pad, dims), _tup(i2h_dilate, dims), _tup(h2h_dilate, dims)
        self._h2h_pad = tuple(d * (k - 1) // 2 for d, k in zip(self._h2h_dilate, self._h2h_kernel))
        self._stride = (1,) * dims
        in_c = self._input_shape[0]
        spatial = self._input_shape[1:]
        self._state_shape = (hidden_channels,) + tuple(
            (s + 2 * p - d * (k - 1) - 1) + 1 for s, p, d, k in zip(spatial, self._i2h_pad, self._i2h_dilate, self._i2h_kernel))
        g = self._ngates
        self.i2h_weight = self.params.get("i2h_weight", shape=(g * hidden_channels, in_c) + self._i2h_kernel, init=i2h_weight_initializer,
                                          allow_deferred_init=True)
        self.h2h_weight = self.params.get("h2h_weight", shape=(g * hidden_channels, hidden_channels) + self._h2h_kernel, init=h2h_weight_initializer)
        self.i2h_bias = self.params.get("i2h_bias", shape=(g * hidden_channels,), init=i2h_bias_initializer)
        self.h2h_bias = self.params.get("h2h_bias", shape=(g * hidden_channels,), init=h2h_bias_initializer)

    def state_info(self, batch_size=0):
        return [{"shape": (batch_size,) + self._state_shape, "__layout__": "NC" + "DHW"[3 - self._dims:]}]

    def _conv_forward(self, F, inputs, states, i2h_weight, h2h_weight, i2h_bias, h2h_bias):
        g = self._ngates
        prefix = "t%d_" % self._counter
        i2h = F.Convolution(inputs, i2h_weight, i2h_bias, kernel=self._i2h_kernel, stride=self._stride, pad=self._i2h_pad,
                            dilate=self._i2h_dilate, num_filter=g * self._hidden_channels, name=prefix + "i2h")
        h2h = F.Convolution(states[0], h2h_weight, h2h_bias, kernel=self._h2h_kernel, stride=self._stride, pad=self._h2h_pad,
                            dilate=self._h2h_dilate, num_filter=g * self._hidden_channels, name=prefix + "h2h")
        return i2h, h2h

    def _act(self, F, x, name=None):
        kw = {"name": name} if (name is not None and F is not nd) else {}
        return F.Activation(x, act_type=self._activation, **kw) if isinstance(self._activation, str) else self._activation(x, **kw)

    def _infer_param_shapes(self, x, *args):
        # input_shape was given with 0 channels: the channel count comes from the first input
        w = self.i2h_weight
        if w.shape is not None and w.shape[1] == 0:
            w.shape = (w.shape[0], x.shape[1]) + tuple(w.shape[2:])


class _ConvRNNCell(_BaseConvRNNCell):
    _ngates = 1

    def _alias(self):
        return "conv_rnn"

    def hybrid_forward(self, F, inputs, states, i2h_weight, h2h_weight, i2h_bias, h2h_bias):
        i2h, h2h = self._conv_forward(F, inputs, states, i2h_weight, h2h_weight, i2h_bias, h2h_bias)
        out = self._act(F, i2h + h2h, name="t%d_out" % self._counter)
        return out, [out]


class _ConvLSTMCell(_BaseConvRNNCell):
    _ngates = 4

    def _alias(self):
        return "conv_lstm"

    def state_info(self, batch_size=0):
        s = super(_ConvLSTMCell, self).state_info(batch_size)[0]
        return [s, dict(s)]

    def hybrid_forward(self, F, inputs, states, i2h_weight, h2h_weight, i2h_bias, h2h_bias):
        i2h, h2h = self._conv_forward(F, inputs, states, i2h_weight, h2h_weight, i2h_bias, h2h_bias)
        s = F.split(i2h + h2h, num_outputs=4, axis=1)
        in_gate, forget_gate = F.sigmoid(s[0]), F.sigmoid(s[1])
        in_transform, out_gate = self._act(F, s[2]), F.sigmoid(s[3])
        prefix = "t%d_" % self._counter
        next_c = F._internal._plus(forget_gate * states[1], in_gate * in_transform, name=prefix + "state")
        next_h = F._internal._mul(out_gate, self._act(F, next_c), name=prefix + "out")
        return next_h, [next_h, next_c]


class _ConvGRUCell(_BaseConvRNNCell):
    _ngates = 3

    def _alias(self):
        return "conv_gru"

    def hybrid_forward(self, F, inputs, states, i2h_weight, h2h_weight, i2h_bias, h2h_bias):
        i2h, h2h = self._conv_forward(F, inputs, states, i2h_weight, h2h_weight, i2h_bias, h2h_bias)
        i2h_r, i2h_z, i2h_n = F.split(i2h, num_outputs=3, axis=1)
        h2h_r, h2h_z, h2h_n = F.split(h2h, num_outputs=3, axis=1)
        reset_gate, update_gate = F.sigmoid(i2h_r + h2h_r), F.sigmoid(i2h_z + h2h_z)
        cand = self._act(F, i2h_n + reset_gate * h2h_n)
        next_h = F._internal._plus((1. - update_gate) * cand, update_gate * states[0], name="t%d_out" % self._counter)
        return next_h, [next_h]


def _make(base, dims, name, doc):
    layout = {1: "NCW", 2: "NCHW", 3: "NCDHW"}[dims]

    def __init__(self, input_shape, hidden_channels, i2h_kernel, h2h_kernel, i2h_pad=(0,) * dims, i2h_dilate=(1,) * dims,
                 h2h_dilate=(1,) * dims, i2h_weight_initializer=None, h2h_weight_initializer=None, i2h_bias_initializer="zeros",
                 h2h_bias_initializer="zeros", conv_layout=layout, activation="tanh", prefix=None, params=None):
        base.__init__(self, input_shape, hidden_channels, i2h_kernel, h2h_kernel, i2h_pad, i2h_dilate, h2h_dilate,
                      i2h_weight_initializer, h2h_weight_initializer, i2h_bias_initializer, h2h_bias_initializer, dims,
                      conv_layout, activation, prefix, params)
    return type(name, (base,), {"__init__": __init__, "__doc__": doc})


Conv1DRNNCell = _make(_ConvRNNCell, 1, "Conv1DRNNCell", "1D convolutional RNN cell (conv_rnn_cell.py:218).")
Conv2DRNNCell = _make(_ConvRNNCell, 2, "Conv2DRNNCell", "2D convolutional RNN cell (conv_rnn_cell.py:285).")
Conv3DRNNCell = _make(_ConvRNNCell, 3, "Conv3DRNNCell", "3D convolutional RNN cell (conv_rnn_cell.py:352).")
Conv1DLSTMCell = _make(_ConvLSTMCell, 1, "Conv1DLSTMCell", "1D ConvLSTM cell (conv_rnn_cell.py:473).")
Conv2DLSTMCell = _make(_ConvLSTMCell, 2, "Conv2DLSTMCell", "2D ConvLSTM cell (conv_rnn_cell.py:550).")
Conv3DLSTMCell = _make(_ConvLSTMCell, 3, "Conv3DLSTMCell", "3D ConvLSTM cell (conv_rnn_cell.py:627).")
Conv1DGRUCell = _make(_ConvGRUCell, 1, "Conv1DGRUCell", "1D ConvGRU cell (conv_rnn_cell.py:762).")
Conv2DGRUCell = _make(_ConvGRUCell, 2, "Conv2DGRUCell", "2D ConvGRU cell (conv_rnn_cell.py:834).")
Conv3DGRUCell = _make(_ConvGRUCell, 3, "Conv3DGRUCell", "3D ConvGRU cell (conv_rnn_cell.py:906).")
