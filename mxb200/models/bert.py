"""BERT encoder (+ MLM / NSP pre-training heads) on the B200 kernels.

The reference tree has no transformer model; its building blocks are FullyConnected, LayerNorm, Embedding/take,
batch_dot, softmax, Dropout and ``_contrib_div_sqrt_dim`` (last table row) and its LAMB-era
recipe is "BERT-base pretraining with gluon.Trainer" (config 4).  Mapping here:

  FullyConnected (+bias)      -> tcgen05 GEMM, bias in the epilogue (ops/functional.py::linear); dgrad / wgrad are the
                                 NN / TN modes of the same kernel, wgrad reduces in fp32 straight into the LAMB arena
  bias + GELU                 -> one fused elementwise kernel each way (csrc/cuda/layernorm.cu::bias_act_*)
  residual add + LayerNorm    -> one kernel (LN(x + r)), fp32 statistics, dgamma/dbeta into the arena
  batch_dot/softmax/batch_dot -> library flash attention through torch SDPA (counts as a library call, like cuBLAS)
  softmax cross entropy       -> fused loss+gradient kernel (csrc/cuda/softmax_ce.cu)

Parameters live in an ordered dict name -> tensor exactly like models/resnet.py so that the same DeviceTrainer
(symmetric arenas, fused allreduce + LAMB update) drives it.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

from ..ops import functional as K
from ..ops.attention import self_attention, advance_rng_step
from ..ops.layernorm import bias_act, layer_norm


class BertSpec(object):
    def __init__(self, vocab=30522, hidden=768, layers=12, heads=12, intermediate=3072, max_pos=512, type_vocab=2,
                 dropout=0.1, eps=1e-12):
        assert hidden % heads == 0
        self.vocab, self.hidden, self.layers, self.heads = vocab, hidden, layers, heads
        self.intermediate, self.max_pos, self.type_vocab = intermediate, max_pos, type_vocab
        self.dropout, self.eps = dropout, eps
        # the MLM decoder is a [vocab_padded, hidden] GEMM: pad the vocabulary to a multiple of 8 (TMA row stride)
        self.vocab_padded = (vocab + 7) // 8 * 8


def bert_base():
    return BertSpec()


def bert_large():
    return BertSpec(hidden=1024, layers=24, heads=16, intermediate=4096)


def init_bert_params(spec, dtype=torch.bfloat16, device="cpu", seed=0):
    """name -> tensor.  Matrices are ``dtype`` (bf16), biases and LayerNorm parameters stay fp32 (they are tiny and the
    kernels take fp32 bias/gamma/beta), init = N(0, 0.02) as in the BERT recipe."""
    gen = torch.Generator().manual_seed(seed)
    P = OrderedDict()

    def mat(name, rows, cols):
        P[name] = (torch.randn((rows, cols), generator=gen) * 0.02).to(device=device, dtype=dtype)

    def vec(name, n, value=0.0):
        P[name] = torch.full((n,), value, dtype=torch.float32, device=device)

    H, I = spec.hidden, spec.intermediate
    mat("word_embed", spec.vocab_padded, H)
    mat("pos_embed", spec.max_pos, H)
    mat("type_embed", spec.type_vocab, H)
    vec("embed_ln_gamma", H, 1.0); vec("embed_ln_beta", H)
    for l in range(spec.layers):
        p = "layer%d_" % l
        mat(p + "qkv_weight", 3 * H, H); vec(p + "qkv_bias", 3 * H)
        mat(p + "proj_weight", H, H); vec(p + "proj_bias", H)
        vec(p + "ln1_gamma", H, 1.0); vec(p + "ln1_beta", H)
        mat(p + "ffn1_weight", I, H); vec(p + "ffn1_bias", I)
        mat(p + "ffn2_weight", H, I); vec(p + "ffn2_bias", H)
        vec(p + "ln2_gamma", H, 1.0); vec(p + "ln2_beta", H)
    mat("pooler_weight", H, H); vec("pooler_bias", H)
    mat("mlm_dense_weight", H, H); vec("mlm_dense_bias", H)
    vec("mlm_ln_gamma", H, 1.0); vec("mlm_ln_beta", H)
    vec("mlm_decoder_bias", spec.vocab_padded)
    mat("nsp_weight", 8, H); vec("nsp_bias", 8)      # 2 classes, padded to 8 output rows for the GEMM
    return P


def _sub(g, lo, hi):
    """Row slice of a gradient-arena sink that keeps the arena tag (ops.functional.sink_touch)."""
    if g is None:
        return None
    v = g[lo:hi]
    ar = getattr(g, "_mxb_arena", None)
    if ar is not None:
        v._mxb_arena = ar
    return v


def _s(G, name):
    return G.get(name) if G is not None else None


def _linear(x, P, G, name, bias=True):
    return K.linear(x, P[name + "_weight"], P[name + "_bias"] if bias else None, sink_w=_s(G, name + "_weight"),
                    sink_b=_s(G, name + "_bias") if bias else None)


def _ln(x, P, G, name, eps, residual=None):
    return layer_norm(x, P[name + "_gamma"], P[name + "_beta"], eps, sink_g=_s(G, name + "_gamma"),
                      sink_b=_s(G, name + "_beta"), residual=residual)


def _dropout(x, p, training):
    return F.dropout(x, p, training) if (training and p > 0) else x


def bert_encoder(spec, P, input_ids, token_types=None, attention_mask=None, training=True, G=None):
    """input_ids [B, S] int64 -> hidden states [B, S, H].  attention_mask: [B, S] (1 = token, 0 = padding) or None."""
    B, S = input_ids.shape
    H, nh = spec.hidden, spec.heads
    pos = torch.arange(S, device=input_ids.device)
    x = F.embedding(input_ids, P["word_embed"]) + P["pos_embed"][pos].unsqueeze(0)
    if token_types is not None:
        x = x + F.embedding(token_types, P["type_embed"])
    else:
        x = x + P["type_embed"][0]
    x = _ln(x, P, G, "embed_ln", spec.eps)
    x = _dropout(x, spec.dropout, training)
    kmask = attention_mask.to(torch.bool) if attention_mask is not None else None
    if training and spec.dropout > 0 and x.is_cuda:
        advance_rng_step(x.device)        # device-side Philox stream offset: fresh masks on every (replayed) step
    for l in range(spec.layers):
        p = "layer%d_" % l
        # Q, K, V as three GEMMs over row-slices of the single [3H, H] parameter: each lands directly in the
        # [B, S, heads, dh] layout the attention kernel reads through strides, and the backward needs no concatenation
        # of dQ / dK / dV (a fused [B, S, 3H] projection costs a 3 x 12.6 MB torch.cat per layer in the backward)
        Wqkv, bqkv = P[p + "qkv_weight"], P[p + "qkv_bias"]
        gw, gb = _s(G, p + "qkv_weight"), _s(G, p + "qkv_bias")
        qkv = []
        for i in range(3):
            qkv.append(K.linear(x, Wqkv[i * H:(i + 1) * H], bqkv[i * H:(i + 1) * H],
                                sink_w=_sub(gw, i * H, (i + 1) * H), sink_b=_sub(gb, i * H, (i + 1) * H)))   # [B, S, H]
        q, k, v = qkv
        # scores, key-padding mask, softmax, attention dropout and the value product in ONE tcgen05 kernel per pass
        # (ops/attention.py -> csrc/cuda/attention.cu); head h lives in columns [64h, 64h+64) of the projections, so no
        # head transposes exist.  Sequences > 128 tokens / other head sizes take the eager composition in there.
        a = self_attention(q, k, v, nh, key_mask=kmask, p_drop=spec.dropout, training=training, call_id=l + 1)
        a = _dropout(_linear(a, P, G, p + "proj"), spec.dropout, training)
        x = _ln(a, P, G, p + "ln1", spec.eps, residual=x)
        h = K.linear(x, P[p + "ffn1_weight"], None, sink_w=_s(G, p + "ffn1_weight"))
        h = _bias_gelu(h, P, G, p + "ffn1_bias")
        h = _dropout(_linear(h, P, G, p + "ffn2"), spec.dropout, training)
        x = _ln(h, P, G, p + "ln2", spec.eps, residual=x)
    return x


def _bias_gelu(h, P, G, bname):
    return bias_act(h, P[bname], "gelu", sink_b=_s(G, bname))


def bert_pretrain_forward(spec, P, input_ids, token_types, attention_mask, masked_positions, training=True, G=None):
    """Returns (mlm_logits [B*Mp, vocab_padded], nsp_logits [B, 8]).  masked_positions: [B, Mp] indices into S."""
    x = bert_encoder(spec, P, input_ids, token_types, attention_mask, training, G)
    B, S, H = x.shape
    idx = (masked_positions + torch.arange(B, device=x.device).unsqueeze(1) * S).reshape(-1)
    m = x.reshape(B * S, H).index_select(0, idx)                           # gather the masked tokens only
    m = K.linear(m, P["mlm_dense_weight"], None, sink_w=_s(G, "mlm_dense_weight"))
    m = _bias_gelu(m, P, G, "mlm_dense_bias")
    m = _ln(m, P, G, "mlm_ln", spec.eps)
    mlm = K.linear(m, P["word_embed"], P["mlm_decoder_bias"], sink_w=_s(G, "word_embed"),
                   sink_b=_s(G, "mlm_decoder_bias"))                        # decoder tied to the word embedding
    pooled = torch.tanh(_linear(x[:, 0], P, G, "pooler"))
    nsp = _linear(pooled, P, G, "nsp")
    return mlm, nsp


def bert_pretrain_loss(spec, mlm_logits, nsp_logits, mlm_labels, nsp_labels):
    """mean MLM cross entropy (over the padded vocabulary; the pad rows never appear as labels) + NSP cross entropy."""
    V = spec.vocab_padded
    if V != spec.vocab:   # keep the padded logits out of the softmax
        pad = torch.zeros(V, dtype=mlm_logits.dtype, device=mlm_logits.device)
        pad[spec.vocab:] = -1e4
        mlm_logits = mlm_logits + pad
    l1 = K.softmax_cross_entropy(mlm_logits, mlm_labels.reshape(-1)).mean()
    nsp_pad = torch.zeros(8, dtype=nsp_logits.dtype, device=nsp_logits.device)
    nsp_pad[2:] = -1e4
    l2 = K.softmax_cross_entropy(nsp_logits + nsp_pad, nsp_labels.reshape(-1)).mean()
    return l1 + l2
