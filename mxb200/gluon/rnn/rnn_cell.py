"""Recurrent cells (reference: python/mxnet/gluon/rnn/rnn_cell.py)."""
from ..block import Block, HybridBlock
from ...ndarray import zeros
from ... import ndarray as nd

__all__ = ["RecurrentCell", "HybridRecurrentCell", "RNNCell", "LSTMCell", "GRUCell", "SequentialRNNCell", "DropoutCell",
           "ResidualCell", "BidirectionalCell", "ZoneoutCell", "HybridSequentialRNNCell", "ModifierCell"]


class RecurrentCell(Block):
    def __init__(self, prefix=None, params=None):
        super(RecurrentCell, self).__init__(prefix=prefix, params=params)
        self._modified = False
        self.reset()

    def reset(self):
        self._init_counter = -1
        self._counter = -1
        for cell in self._children.values():
            cell.reset()

    def state_info(self, batch_size=0):
        raise NotImplementedError()

    def begin_state(self, batch_size=0, func=zeros, **kwargs):
        states = []
        for info in self.state_info(batch_size):
            self._init_counter += 1
            info = dict(info) if info is not None else {}
            info.pop("__layout__", None)
            info.update(kwargs)
            states.append(func(**info))
        return states

    def unroll(self, length, inputs, begin_state=None, layout="NTC", merge_outputs=None, valid_length=None):
        self.reset()
        axis = layout.find("T")
        from ... import symbol as _sym
        if isinstance(inputs, _sym.Symbol) or (isinstance(inputs, (list, tuple)) and inputs and isinstance(inputs[0], _sym.Symbol)):
            return self._unroll_symbolic(_sym, length, inputs, begin_state, axis, merge_outputs, valid_length)
        if isinstance(inputs, nd.NDArray):
            batch_size = inputs.shape[layout.find("N")]
            seq = [x.squeeze(axis) for x in nd.split(inputs, num_outputs=length, axis=axis, squeeze_axis=False)] if length > 1 \
                else [inputs.squeeze(axis)]
        else:
            seq = list(inputs)
            batch_size = seq[0].shape[0]
        if begin_state is None:
            begin_state = self.begin_state(batch_size, ctx=seq[0].context, dtype=seq[0].dtype)
        states = begin_state
        outputs, all_states = [], []
        for i in range(length):
            output, states = self(seq[i], states)
            outputs.append(output)
            if valid_length is not None:
                all_states.append(states)
        if valid_length is not None:
            # padded steps: outputs beyond a sequence's length are zero, the returned states are those of its last valid step
            # (rnn_cell.py:_mask_sequence_variable_length / SequenceLast in the reference)
            states = [nd.SequenceLast(nd.stack(*[st[k] for st in all_states], axis=0), sequence_length=valid_length,
                                      use_sequence_length=True, axis=0) for k in range(len(states))]
            masked = nd.SequenceMask(nd.stack(*outputs, axis=0), sequence_length=valid_length, use_sequence_length=True, axis=0)
            outputs = [masked[i] for i in range(length)]
        merged_in = isinstance(inputs, nd.NDArray)
        if merge_outputs or (merge_outputs is None and merged_in):
            outputs = nd.stack(*outputs, axis=axis)
        return outputs, states

    def _unroll_symbolic(self, F, length, inputs, begin_state, axis, merge_outputs, valid_length=None):
        """unroll() on Symbols (rnn_cell.py:RecurrentCell.unroll of the reference): the steps become graph nodes named
        ``<prefix>t<k>_*``.  Default begin states are zeros derived from the first step's input, so no unknown (0)
        dimension enters the graph."""
        if isinstance(inputs, F.Symbol):
            seq = list(F.split(inputs, axis=axis, num_outputs=length, squeeze_axis=True)) if length > 1 \
                else [F.squeeze(inputs, axis=axis)]
        else:
            seq = list(inputs)
        assert len(seq) == length, "unroll: got %d step inputs for length %d" % (len(seq), length)
        if begin_state is None:
            begin_state = []
            col = F.sum(F.zeros_like(F.reshape(seq[0], shape=(0, -1))), axis=1, keepdims=True)      # (N, 1) of zeros
            for info in self.state_info(0):
                tail = tuple(info["shape"][1:])
                st = F.reshape(col, shape=(-1,) + (1,) * len(tail)) if len(tail) != 1 else col
                for ax, size in enumerate(tail, start=1):
                    st = F.broadcast_axis(st, axis=ax, size=size)
                begin_state.append(st)
        states = begin_state
        outputs, all_states = [], []
        for i in range(length):
            output, states = self(seq[i], states)
            outputs.append(output)
            if valid_length is not None:
                all_states.append(states)
        if valid_length is not None:
            states = [F.SequenceLast(F.stack(*[st[k] for st in all_states], axis=0), sequence_length=valid_length,
                                     use_sequence_length=True, axis=0) for k in range(len(states))]
            masked = F.SequenceMask(F.stack(*outputs, axis=0), sequence_length=valid_length, use_sequence_length=True, axis=0)
            outputs = list(F.split(masked, axis=0, num_outputs=length, squeeze_axis=True)) if length > 1 else [F.squeeze(masked, axis=0)]
        if merge_outputs or (merge_outputs is None and isinstance(inputs, F.Symbol)):
            outputs = F.stack(*outputs, axis=axis)
        return outputs, states

    def forward(self, inputs, states):
        self._counter += 1
        return super(RecurrentCell, self).forward(inputs, states)


class HybridRecurrentCell(RecurrentCell, HybridBlock):
    def __init__(self, prefix=None, params=None):
        super(HybridRecurrentCell, self).__init__(prefix=prefix, params=params)

    def forward(self, inputs, states):
        self._counter += 1
        from ... import symbol as _sym
        if isinstance(inputs, _sym.Symbol):
            params = {i: j.var() for i, j in self._reg_params.items()}
            with self.name_scope():
                return self.hybrid_forward(_sym, inputs, states, **params)
        ctx = inputs.context
        try:
            params = {i: j.data(ctx) for i, j in self._reg_params.items()}
        except Exception:
            self._deferred_infer(inputs, states)
            params = {i: j.data(ctx) for i, j in self._reg_params.items()}
        return self.hybrid_forward(nd, inputs, states, **params)

    def hybrid_forward(self, F, x, *args, **kwargs):
        raise NotImplementedError


class _BaseCell(HybridRecurrentCell):
    _ngates = 1

    def __init__(self, hidden_size, i2h_weight_initializer=None, h2h_weight_initializer=None,
                 i2h_bias_initializer="zeros", h2h_bias_initializer="zeros", input_size=0, prefix=None, params=None):
        super(_BaseCell, self).__init__(prefix=prefix, params=params)
        g = self._ngates
        self._hidden_size, self._input_size = hidden_size, input_size
        self.i2h_weight = self.params.get("i2h_weight", shape=(g * hidden_size, input_size), init=i2h_weight_initializer, allow_deferred_init=True)
        self.h2h_weight = self.params.get("h2h_weight", shape=(g * hidden_size, hidden_size), init=h2h_weight_initializer, allow_deferred_init=True)
        self.i2h_bias = self.params.get("i2h_bias", shape=(g * hidden_size,), init=i2h_bias_initializer, allow_deferred_init=True)
        self.h2h_bias = self.params.get("h2h_bias", shape=(g * hidden_size,), init=h2h_bias_initializer, allow_deferred_init=True)

    def _infer_param_shapes(self, x, *args):
        self.i2h_weight.shape = (self._ngates * self._hidden_size, x.shape[-1])

    def state_info(self, batch_size=0):
        return [{"shape": (batch_size, self._hidden_size), "__layout__": "NC"}]


class RNNCell(_BaseCell):
    def __init__(self, hidden_size, activation="tanh", **kwargs):
        super(RNNCell, self).__init__(hidden_size, **kwargs)
        self._activation = activation

    def _alias(self):
        return "rnn"

    def hybrid_forward(self, F, inputs, states, i2h_weight, h2h_weight, i2h_bias, h2h_bias):
        prefix = "t%d_" % self._counter          # node names of the reference (rnn_cell.py:RNNCell.hybrid_forward)
        i2h = F.FullyConnected(inputs, i2h_weight, i2h_bias, num_hidden=self._hidden_size, name=prefix + "i2h")
        h2h = F.FullyConnected(states[0], h2h_weight, h2h_bias, num_hidden=self._hidden_size, name=prefix + "h2h")
        output = F.Activation(i2h + h2h, act_type=self._activation, name=prefix + "out")
        return output, [output]


class LSTMCell(_BaseCell):
    """``activation`` (cell transform / output squashing, default tanh) and ``recurrent_activation`` (the three gates,
    default sigmoid) take any Activation type name (rnn_cell.py:LSTMCell of the reference)."""
    _ngates = 4

    def __init__(self, hidden_size, activation="tanh", recurrent_activation="sigmoid", **kwargs):
        super(LSTMCell, self).__init__(hidden_size, **kwargs)
        self._activation, self._recurrent_activation = activation, recurrent_activation

    def _alias(self):
        return "lstm"

    def state_info(self, batch_size=0):
        return [{"shape": (batch_size, self._hidden_size), "__layout__": "NC"},
                {"shape": (batch_size, self._hidden_size), "__layout__": "NC"}]

    def hybrid_forward(self, F, inputs, states, i2h_weight, h2h_weight, i2h_bias, h2h_bias):
        prefix = "t%d_" % self._counter
        i2h = F.FullyConnected(inputs, i2h_weight, i2h_bias, num_hidden=self._hidden_size * 4, name=prefix + "i2h")
        h2h = F.FullyConnected(states[0], h2h_weight, h2h_bias, num_hidden=self._hidden_size * 4, name=prefix + "h2h")
        gates = i2h + h2h
        s = F.split(gates, num_outputs=4, axis=1, name=prefix + "slice")
        act = lambda x, kind: F.Activation(x, act_type=kind)   # noqa: E731
        ra, ca = self._recurrent_activation, self._activation
        in_gate, forget_gate = act(s[0], ra), act(s[1], ra)
        in_transform, out_gate = act(s[2], ca), act(s[3], ra)
        next_c = F._internal._plus(forget_gate * states[1], in_gate * in_transform, name=prefix + "state")
        next_h = F._internal._mul(out_gate, act(next_c, ca), name=prefix + "out")
        return next_h, [next_h, next_c]


class GRUCell(_BaseCell):
    _ngates = 3

    def _alias(self):
        return "gru"

    def hybrid_forward(self, F, inputs, states, i2h_weight, h2h_weight, i2h_bias, h2h_bias):
        prefix = "t%d_" % self._counter
        prev = states[0]
        i2h = F.FullyConnected(inputs, i2h_weight, i2h_bias, num_hidden=self._hidden_size * 3, name=prefix + "i2h")
        h2h = F.FullyConnected(prev, h2h_weight, h2h_bias, num_hidden=self._hidden_size * 3, name=prefix + "h2h")
        i2h_r, i2h_z, i2h = F.split(i2h, num_outputs=3, axis=1, name=prefix + "i2h_slice")
        h2h_r, h2h_z, h2h = F.split(h2h, num_outputs=3, axis=1, name=prefix + "h2h_slice")
        reset_gate = F.sigmoid(i2h_r + h2h_r)
        update_gate = F.sigmoid(i2h_z + h2h_z)
        next_h_tmp = F.tanh(i2h + reset_gate * h2h)
        next_h = F._internal._plus((1. - update_gate) * next_h_tmp, update_gate * prev, name=prefix + "out")
        return next_h, [next_h]


class SequentialRNNCell(RecurrentCell):
    def add(self, cell):
        self.register_child(cell)

    def state_info(self, batch_size=0):
        return sum([c.state_info(batch_size) for c in self._children.values()], [])

    def begin_state(self, batch_size=0, **kwargs):
        return sum([c.begin_state(batch_size, **kwargs) for c in self._children.values()], [])

    def __call__(self, inputs, states):
        self._counter += 1
        next_states = []
        p = 0
        for cell in self._children.values():
            n = len(cell.state_info())
            state = states[p:p + n]
            p += n
            inputs, state = cell(inputs, state)
            next_states.append(state)
        return inputs, sum(next_states, [])

    def unroll(self, length, inputs, begin_state=None, layout="NTC", merge_outputs=None, valid_length=None):
        """Layer by layer: every child unrolls over the whole sequence and feeds the next one (rnn_cell.py:
        SequentialRNNCell.unroll of the reference), so children that cannot be stepped (BidirectionalCell) are fine."""
        self.reset()
        children = list(self._children.values())
        if not children:
            return inputs, []
        p = 0
        next_states = []
        for i, cell in enumerate(children):
            n = len(cell.state_info())
            st = None if begin_state is None else list(begin_state[p:p + n])
            p += n
            inputs, st = cell.unroll(length, inputs, begin_state=st, layout=layout,
                                     merge_outputs=None if i < len(children) - 1 else merge_outputs, valid_length=valid_length)
            next_states.extend(st)
        return inputs, next_states

    def __len__(self):
        return len(self._children)

    def __getitem__(self, i):
        return list(self._children.values())[i]


class DropoutCell(HybridRecurrentCell):
    def __init__(self, rate, prefix=None, params=None):
        super(DropoutCell, self).__init__(prefix, params)
        self.rate = rate

    def state_info(self, batch_size=0):
        return []

    def _alias(self):
        return "dropout"

    def hybrid_forward(self, F, inputs, states):
        if self.rate > 0:
            inputs = F.Dropout(inputs, p=self.rate)
        return inputs, states


class ResidualCell(HybridRecurrentCell):
    def __init__(self, base_cell):
        super(ResidualCell, self).__init__(prefix=base_cell.prefix + "res_", params=None)
        self.base_cell = base_cell

    def state_info(self, batch_size=0):
        return self.base_cell.state_info(batch_size)

    def begin_state(self, batch_size=0, **kwargs):
        return self.base_cell.begin_state(batch_size, **kwargs)

    def __call__(self, inputs, states):
        output, states = self.base_cell(inputs, states)
        return output + inputs, states

    def unroll(self, length, inputs, begin_state=None, layout="NTC", merge_outputs=None, valid_length=None):
        """Unrolls the base cell as a whole (so cells that cannot be stepped, like BidirectionalCell, work) and adds the
        inputs step by step (rnn_cell.py:ResidualCell.unroll of the reference)."""
        self.reset()
        self.base_cell._modified = False
        outputs, states = self.base_cell.unroll(length, inputs, begin_state=begin_state, layout=layout, merge_outputs=merge_outputs,
                                                valid_length=valid_length)
        self.base_cell._modified = True
        axis = layout.find("T")
        F, seq_in = _as_steps(inputs, length, axis)
        if isinstance(outputs, (list, tuple)):
            outputs = [o + i for o, i in zip(outputs, seq_in)]
            if valid_length is not None:      # the residual re-introduces the padded inputs: mask again
                m = F.SequenceMask(F.stack(*outputs, axis=0), sequence_length=valid_length, use_sequence_length=True, axis=0)
                outputs = [m[i] for i in range(length)] if F is nd else list(F.split(m, axis=0, num_outputs=length, squeeze_axis=True))
        else:
            outputs = outputs + (F.stack(*seq_in, axis=axis) if isinstance(inputs, (list, tuple)) else inputs)
            if valid_length is not None:
                outputs = F.SequenceMask(outputs, sequence_length=valid_length, use_sequence_length=True, axis=axis)
        return outputs, states


def _as_steps(inputs, length, axis):
    """(namespace, [step inputs]) for a merged tensor / symbol or a list of per-step tensors / symbols."""
    from ... import symbol as _sym
    first = inputs[0] if isinstance(inputs, (list, tuple)) else inputs
    F = _sym if isinstance(first, _sym.Symbol) else nd
    if isinstance(inputs, (list, tuple)):
        seq = list(inputs)
    elif length > 1:
        seq = list(F.split(inputs, axis=axis, num_outputs=length, squeeze_axis=True))
    else:
        seq = [F.squeeze(inputs, axis=axis)]
    assert len(seq) == length, "unroll: got %d step inputs for length %d" % (len(seq), length)
    return F, seq


class ZoneoutCell(HybridRecurrentCell):
    """Zoneout (Krueger et al. 2016): with probability ``zoneout_states`` a state unit keeps its previous value, with
    ``zoneout_outputs`` an output unit keeps the previous output (rnn_cell.py:ZoneoutCell of the reference)."""

    def __init__(self, base_cell, zoneout_outputs=0., zoneout_states=0.):
        assert not isinstance(base_cell, BidirectionalCell), "BidirectionalCell doesn't support zoneout since it doesn't " \
            "support step. Please add ZoneoutCell to the cells underneath instead."
        super(ZoneoutCell, self).__init__(prefix=base_cell.prefix + "zoneout_", params=None)
        self.base_cell = base_cell
        self.zoneout_outputs, self.zoneout_states = zoneout_outputs, zoneout_states
        self._prev_output = None

    def state_info(self, batch_size=0):
        return self.base_cell.state_info(batch_size)

    def begin_state(self, batch_size=0, **kwargs):
        return self.base_cell.begin_state(batch_size, **kwargs)

    def _alias(self):
        return "zoneout"

    def reset(self):
        super(ZoneoutCell, self).reset()
        if getattr(self, "base_cell", None) is not None:      # reset() is also called by the base constructor
            self.base_cell.reset()
        self._prev_output = None

    def __call__(self, inputs, states):
        from ... import symbol as _sym
        F = _sym if isinstance(inputs, _sym.Symbol) else nd
        next_output, next_states = self.base_cell(inputs, states)

        def mask(p, like):
            return F.Dropout(F.ones_like(like), p=p)
        prev_output = self._prev_output if self._prev_output is not None else F.zeros_like(next_output)
        output = F.where(mask(self.zoneout_outputs, next_output), next_output, prev_output) \
            if self.zoneout_outputs != 0. else next_output
        new_states = [F.where(mask(self.zoneout_states, ns), ns, os) for ns, os in zip(next_states, states)] \
            if self.zoneout_states != 0. else next_states
        self._prev_output = output
        return output, new_states


class BidirectionalCell(HybridRecurrentCell):
    def __init__(self, l_cell, r_cell, output_prefix="bi_"):
        super(BidirectionalCell, self).__init__(prefix="", params=None)
        self.l_cell, self.r_cell = l_cell, r_cell
        self._output_prefix = output_prefix

    def state_info(self, batch_size=0):
        return self.l_cell.state_info(batch_size) + self.r_cell.state_info(batch_size)

    def begin_state(self, batch_size=0, **kwargs):
        return self.l_cell.begin_state(batch_size, **kwargs) + self.r_cell.begin_state(batch_size, **kwargs)

    def __call__(self, inputs, states):
        raise NotImplementedError("Bidirectional cannot be stepped. Please use unroll")

    def unroll(self, length, inputs, begin_state=None, layout="NTC", merge_outputs=None, valid_length=None):
        """Left cell over the sequence, right cell over the reversed sequence, per-step concatenation named
        ``<output_prefix>t<k>_output`` (rnn_cell.py:BidirectionalCell.unroll of the reference).  Inputs: a merged NDArray /
        Symbol or a list of per-step ones."""
        self.reset()
        axis = layout.find("T")
        F, seq = _as_steps(inputs, length, axis)
        nl = len(self.l_cell.state_info())
        if begin_state is None:
            l_begin = r_begin = None
        else:
            l_begin, r_begin = list(begin_state[:nl]), list(begin_state[nl:])
        if valid_length is None:
            rev_in = list(reversed(seq))
        else:       # reverse every sequence inside its own length (padding stays at the end)
            rv = F.SequenceReverse(F.stack(*seq, axis=0), sequence_length=valid_length, use_sequence_length=True, axis=0)
            rev_in = [rv[i] for i in range(length)] if F is nd else list(F.split(rv, axis=0, num_outputs=length, squeeze_axis=True))
        l_out, l_st = self.l_cell.unroll(length, seq, begin_state=l_begin, layout=layout, merge_outputs=False, valid_length=valid_length)
        r_out, r_st = self.r_cell.unroll(length, rev_in, begin_state=r_begin, layout=layout, merge_outputs=False, valid_length=valid_length)
        if valid_length is None:
            r_out = list(reversed(list(r_out)))
        else:
            rr = F.SequenceReverse(F.stack(*r_out, axis=0), sequence_length=valid_length, use_sequence_length=True, axis=0)
            r_out = [rr[i] for i in range(length)] if F is nd else list(F.split(rr, axis=0, num_outputs=length, squeeze_axis=True))
        if F is nd:
            outs = [nd.concat(lo, ro, dim=1) for lo, ro in zip(l_out, r_out)]
        else:
            outs = [F.Concat(lo, ro, dim=1, name="%st%d" % (self._output_prefix, i)) for i, (lo, ro) in enumerate(zip(l_out, r_out))]
        merged_in = not isinstance(inputs, (list, tuple))
        if merge_outputs or (merge_outputs is None and merged_in):
            outs = F.stack(*outs, axis=axis)
        return outs, list(l_st) + list(r_st)


# the reference distinguishes a hybridizable stack; every cell here runs on both NDArrays and Symbols
class HybridSequentialRNNCell(SequentialRNNCell, HybridBlock):
    """SequentialRNNCell that may be nested inside a HybridBlock (its children must be hybrid cells themselves)."""

    def __init__(self, prefix=None, params=None):
        super(HybridSequentialRNNCell, self).__init__(prefix=prefix, params=params)

ModifierCell = ResidualCell.__mro__[1] if False else RecurrentCell
