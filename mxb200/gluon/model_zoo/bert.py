"""Gluon front-end of the BERT encoder / pre-training model (mxb200/models/bert.py): parameters are ordinary
``gluon.Parameter`` objects (so ``collect_params``, ``save_parameters``, ``gluon.Trainer`` with the fused LAMB
arena path all work), the forward pass is the functional model on the B200 kernels."""
import torch

from ...models import bert as _bert
from ...ndarray.ndarray import NDArray
from ..block import HybridBlock

__all__ = ["BERTModel", "bert_12_768_12", "bert_24_1024_16", "get_bert"]


class BERTModel(HybridBlock):
    """BERT with MLM + NSP heads.  ``forward(input_ids, token_types, valid_mask, masked_positions)`` returns
    (mlm_logits [B*Mp, vocab_padded], nsp_logits [B, 8]); with ``masked_positions=None`` the encoder output."""

    def __init__(self, spec=None, dtype="bfloat16", seed=0, prefix=None, params=None, **spec_kw):
        super(BERTModel, self).__init__(prefix=prefix, params=params)
        self.spec = spec or _bert.BertSpec(**spec_kw)
        tdtype = {"bfloat16": torch.bfloat16, "float32": torch.float32, "float16": torch.float16}[str(dtype)]
        init = _bert.init_bert_params(self.spec, dtype=tdtype, device="cpu", seed=seed)
        self._pnames = list(init.keys())
        from ... import initializer as _init
        with self.name_scope():
            for n, t in init.items():
                p = self.params.get(n, shape=tuple(t.shape), dtype={torch.bfloat16: "bfloat16", torch.float32: "float32",
                                                                    torch.float16: "float16"}[t.dtype],
                                    init=_init.Constant(0.0))
                p._bert_init = t
                setattr(self, "_p_" + n, p)

    def initialize(self, init=None, ctx=None, verbose=False, force_reinit=False):
        super(BERTModel, self).initialize(init=init, ctx=ctx, verbose=verbose, force_reinit=force_reinit)
        for n in self._pnames:
            p = getattr(self, "_p_" + n)
            for d in p.list_data():
                with torch.no_grad():
                    d._t.copy_(p._bert_init.to(d._t.device))
        return self

    def _P(self):
        out, G = {}, {}
        for n in self._pnames:
            p = getattr(self, "_p_" + n)
            out[n] = p.data()._t
            sink = getattr(p, "_grad_sink", None)
            if sink is not None:
                G[n] = sink
        return out, (G or None)

    def forward(self, input_ids, token_types=None, valid_mask=None, masked_positions=None):
        from ... import autograd
        P, G = self._P()
        tt = token_types._t.long() if token_types is not None else None
        am = valid_mask._t if valid_mask is not None else None
        ids = input_ids._t.long()
        training = autograd.is_training()
        if masked_positions is None:
            return NDArray(_bert.bert_encoder(self.spec, P, ids, tt, am, training, G))
        mlm, nsp = _bert.bert_pretrain_forward(self.spec, P, ids, tt, am, masked_positions._t.long(), training, G)
        return NDArray(mlm), NDArray(nsp)


def get_bert(name="bert_12_768_12", **kw):
    spec = {"bert_12_768_12": _bert.bert_base, "bert_24_1024_16": _bert.bert_large}[name]()
    return BERTModel(spec=spec, **kw)


def bert_12_768_12(**kw):
    return get_bert("bert_12_768_12", **kw)


def bert_24_1024_16(**kw):
    return get_bert("bert_24_1024_16", **kw)
