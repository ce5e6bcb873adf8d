"""contrib.quantization: post-training int8 calibration helpers (reference: python/mxnet/contrib/quantization.py).

``quantize_model`` in the reference rewrites the symbol with quantized operators (a C++ graph pass) and calibrates
thresholds with either the observed min/max ('naive') or the KL-divergence optimal threshold ('entropy').  Here the
calibration machinery is provided (layer-output collection through a Module, min/max and entropy thresholds) and
``quantize_model`` returns the ORIGINAL symbol with int8-quantised parameter tensors + their ranges and the calibration
table: B200 int8 tensor-core inference kernels are not part of this framework (bf16 is the production dtype)."""
import logging

import numpy as np

from .. import ndarray as nd


def _smooth_distribution(p, eps=0.0001):
    is_zeros = (p == 0).astype(np.float32)
    is_nonzeros = (p != 0).astype(np.float32)
    n_zeros, n_nonzeros = is_zeros.sum(), p.size - is_zeros.sum()
    if not n_nonzeros:
        raise ValueError("The discrete probability distribution is malformed. All entries are 0.")
    eps1 = eps * float(n_zeros) / float(n_nonzeros)
    assert eps1 < 1.0, "n_zeros=%d, n_nonzeros=%d, eps1=%f" % (n_zeros, n_nonzeros, eps1)
    hist = p.astype(np.float32)
    hist += eps * is_zeros + (-eps1) * is_nonzeros
    return hist


def _get_optimal_threshold(arr, num_bins=8001, num_quantized_bins=255):
    """KL-divergence calibration (quantization.py:266-353): the clipping threshold whose quantised distribution is closest
    to the observed one.  Returns (min, max, optimal_divergence_threshold_min?, threshold)."""
    arr = arr.asnumpy() if isinstance(arr, nd.NDArray) else np.asarray(arr)
    min_val, max_val = float(arr.min()), float(arr.max())
    th = max(abs(min_val), abs(max_val))
    if th == 0:
        return min_val, max_val, 0.0, 0.0
    hist, edges = np.histogram(arr, bins=num_bins, range=(-th, th))
    zero = num_bins // 2
    half_q = num_quantized_bins // 2
    best_div, best_th = np.inf, th
    for i in range(half_q, zero + 1, max(1, (zero - half_q) // 128)):
        lo, hi = zero - i, zero + i + 1
        sliced = hist[lo:hi].astype(np.float64)
        p = sliced.copy()
        p[0] += hist[:lo].sum()
        p[-1] += hist[hi:].sum()
        nz = (sliced != 0)
        factor = len(sliced) // num_quantized_bins
        q = np.zeros_like(sliced)
        for j in range(num_quantized_bins):
            s, e = j * factor, (j + 1) * factor if j != num_quantized_bins - 1 else len(sliced)
            cnt = nz[s:e].sum()
            if cnt:
                q[s:e] = np.where(nz[s:e], sliced[s:e].sum() / cnt, 0)
        try:
            ps, qs = _smooth_distribution(p), _smooth_distribution(q)
        except ValueError:
            continue
        ps, qs = ps / ps.sum(), qs / qs.sum()
        div = float(np.sum(np.where(ps > 0, ps * np.log(ps / np.maximum(qs, 1e-12)), 0)))
        if div < best_div:
            best_div, best_th = div, float(edges[hi] if hi < len(edges) else th)
    return min_val, max_val, best_div, best_th


def _collect_layer_outputs(mod, data, max_num_examples=None, include=None):
    """name -> list of output arrays of every internal layer (Module with all internals bound)."""
    outs = {}
    data.reset()
    seen = 0
    for batch in data:
        mod.forward(batch, is_train=False)
        for name, arr in zip(mod.output_names, mod.get_outputs()):
            if include is None or include(name):
                outs.setdefault(name, []).append(arr.asnumpy())
        seen += batch.data[0].shape[0]
        if max_num_examples is not None and seen >= max_num_examples:
            break
    return outs, seen


def calibrate(sym, arg_params, aux_params, calib_data, ctx=None, calib_mode="naive", num_calib_examples=None,
              data_names=("data",), label_names=("softmax_label",), logger=logging):
    """layer-output name -> (min, max) thresholds ('naive': observed range, 'entropy': KL-optimal symmetric range)."""
    from ..context import cpu
    from ..module import Module
    internals = sym.get_internals()
    mod = Module(internals, data_names=data_names, label_names=None, context=ctx or cpu())
    mod.bind(for_training=False, data_shapes=calib_data.provide_data)
    mod.set_params(arg_params, aux_params, allow_missing=True, allow_extra=True)
    outs, n = _collect_layer_outputs(mod, calib_data, num_calib_examples, include=lambda s: s.endswith("_output"))
    logger.info("Collected layer outputs from %d examples", n)
    table = {}
    for name, arrs in outs.items():
        a = np.concatenate([x.reshape(-1) for x in arrs])
        if calib_mode == "entropy":
            _mn, _mx, _d, th = _get_optimal_threshold(a)
            table[name] = (-th, th)
        else:
            table[name] = (float(a.min()), float(a.max()))
    return table


def quantize_params(arg_params):
    """weight/bias tensors -> {name_quantize: int8 NDArray, name_min/_max: ranges} (quantization.py:43-80)."""
    out = {}
    for k, v in arg_params.items():
        if not k.endswith(("weight", "bias")):
            out[k] = v
            continue
        lo, hi = nd.min(v), nd.max(v)
        q, qlo, qhi = nd.contrib.quantize(v, lo.reshape((1,)), hi.reshape((1,)), out_type="int8")
        out[k + "_quantize"], out[k + "_min"], out[k + "_max"] = q, qlo, qhi
    return out


def quantize_model(sym, arg_params, aux_params, data_names=("data",), label_names=("softmax_label",), ctx=None,
                   excluded_sym_names=None, calib_mode="entropy", calib_data=None, num_calib_examples=None,
                   calib_layer=None, quantized_dtype="int8", logger=logging):
    """Returns (symbol, quantised arg params, aux params[, calibration table as attribute ``calib_table``])."""
    if quantized_dtype not in ("int8", "uint8"):
        raise ValueError("unknown quantized_dtype %s received, expected `int8` or `uint8`" % quantized_dtype)
    table = {}
    if calib_mode is not None and calib_mode != "none":
        if calib_data is None:
            raise ValueError("calib_data must be provided when calib_mode=%s" % calib_mode)
        table = calibrate(sym, arg_params, aux_params, calib_data, ctx, calib_mode, num_calib_examples, data_names, label_names, logger)
    qargs = quantize_params(arg_params)
    quantize_model.calib_table = table
    return sym, qargs, aux_params
