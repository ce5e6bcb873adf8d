"""Python launchers for the fused optimizer kernels (csrc/cuda/optim.cu)."""
import ctypes

import torch

from .. import _native as N
from ._dtypes import code

c_void_p, c_int, c_int64, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

LAUNCHES = [0]  # kernels of ours launched (bench.py reports this)


def _p(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def sgd_update(w, g, mom, w32, lr, wd, momentum, rescale, clip, nesterov=False, chunk_tid=None, lr_mult=None,
               wd_mult=None, zero_grad=False):
    """In-place fused SGD / momentum / NAG (+fp32 master).  ``w`` may be None when w32 is the only weight copy."""
    lib = N.cuda()
    ref = w if w is not None else w32
    n = ref.numel()
    clip = -1.0 if clip is None else float(clip)
    rc = lib.MXBSgdUpdate(_p(w), c_int(code(w.dtype) if w is not None else 0), _p(g), c_int(code(g.dtype)), _p(mom),
                          c_int(code(mom.dtype) if mom is not None else 0), _p(w32), c_int64(n), _p(chunk_tid),
                          _p(lr_mult), _p(wd_mult), c_float(lr), c_float(wd), c_float(momentum), c_float(rescale),
                          c_float(clip), c_int(1 if nesterov else 0), c_int(1 if zero_grad else 0), N.stream_of(ref))
    N.check(lib, rc)
    LAUNCHES[0] += 1


def adam_update(w, g, m, v, w32, lr, beta1, beta2, eps, wd, rescale, clip, chunk_tid=None, lr_mult=None, wd_mult=None,
                zero_grad=False):
    lib = N.cuda()
    ref = w if w is not None else w32
    clip = -1.0 if clip is None else float(clip)
    rc = lib.MXBAdamUpdate(_p(w), c_int(code(w.dtype) if w is not None else 0), _p(g), c_int(code(g.dtype)), _p(m), _p(v),
                           _p(w32), c_int64(ref.numel()), _p(chunk_tid), _p(lr_mult), _p(wd_mult), c_float(lr),
                           c_float(wd), c_float(beta1), c_float(beta2), c_float(eps), c_float(rescale), c_float(clip),
                           c_int(1 if zero_grad else 0), N.stream_of(ref))
    N.check(lib, rc)
    LAUNCHES[0] += 1


def lamb_phase1(w, g, m, v, w32, r, w_norm2, r_norm2, beta1, beta2, eps, rescale, clip, bc1, bc2, wd, chunk_tid=None,
                wd_mult=None, zero_grad=False):
    lib = N.cuda()
    ref = w if w is not None else w32
    clip = -1.0 if clip is None else float(clip)
    rc = lib.MXBLambPhase1(_p(w), c_int(code(w.dtype) if w is not None else 0), _p(g), c_int(code(g.dtype)), _p(m), _p(v),
                           _p(w32), _p(r), c_int64(ref.numel()), _p(chunk_tid), _p(wd_mult), _p(w_norm2), _p(r_norm2),
                           c_float(beta1), c_float(beta2), c_float(eps), c_float(rescale), c_float(clip), c_float(bc1),
                           c_float(bc2), c_float(wd), c_int(1 if zero_grad else 0), N.stream_of(ref))
    N.check(lib, rc)
    LAUNCHES[0] += 1


def lamb_phase2(w, w32, r, w_norm2, r_norm2, lr, lower=-1.0, upper=-1.0, chunk_tid=None, lr_mult=None):
    lib = N.cuda()
    ref = w if w is not None else w32
    rc = lib.MXBLambPhase2(_p(w), c_int(code(w.dtype) if w is not None else 0), _p(w32), _p(r), c_int64(ref.numel()),
                           _p(chunk_tid), _p(lr_mult), _p(w_norm2), _p(r_norm2), c_float(lr), c_float(lower),
                           c_float(upper), N.stream_of(ref))
    N.check(lib, rc)
    LAUNCHES[0] += 1


def elementwise_sum(out, ins, scale=1.0):
    """out = scale * sum(ins) (fp32 accumulate), <= 8 inputs per launch; more inputs are chained."""
    lib = N.cuda()
    ins = list(ins)
    first = True
    while ins:
        batch, ins = ins[:8 if first else 7], ins[8 if first else 7:]
        srcs = batch if first else [out] + batch
        arr = (c_void_p * len(srcs))(*[t.data_ptr() for t in srcs])
        rc = lib.MXBElementwiseSum(_p(out), c_int(code(out.dtype)), arr, c_int(len(srcs)), c_int64(out.numel()),
                                   c_float(scale if not ins else 1.0), N.stream_of(out))
        N.check(lib, rc)
        LAUNCHES[0] += 1
        first = False
    return out
