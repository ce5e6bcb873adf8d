// Per-element optimizer math shared by the stand-alone update kernels (optim.cu) and the fused
// allreduce+update kernels (allreduce.cu).  Formulas and sign conventions follow the reference:
//   SGD       : w = (1-lr*wd)*w - lr*clip(rescale*g)                          optimizer_op-inl.h:86-125
//   SGD-mom   : m = mu*m - lr*wd*w - lr*clip(rescale*g);  w += m               optimizer_op-inl.h:305-353
//   NAG       : g' = clip(rescale*g) + wd*w; m = mu*m + g'; w -= lr*(g' + mu*m) python/mxnet/optimizer/optimizer.py:952-973
//   Adam      : g' = clip(rescale*g + wd*w); m,v EMA; w -= lr*m/(sqrt(v)+eps)  optimizer_op-inl.h:841-876
//   LAMB      : not in the reference; You et al. 2019, two-phase with per-tensor trust ratio.
#pragma once
#include "common.cuh"

namespace mxb {

enum OptKind : int { kOptSGD = 0, kOptNAG = 1 };

struct SgdHyper {
  float lr, wd, momentum, rescale, clip;  // clip < 0 disables
  int nesterov;
  int has_mom;
};

__device__ __forceinline__ float clip_grad(float g, float clip) {
  return clip >= 0.f ? fminf(fmaxf(g, -clip), clip) : g;
}

// Updates w (fp32 master value) and m in registers.
__device__ __forceinline__ void sgd_step(float& w, float g, float& m, const SgdHyper& h, float lr, float wd) {
  g = clip_grad(g * h.rescale, h.clip);
  if (!h.has_mom) {
    w = (1.f - lr * wd) * w - lr * g;
  } else if (!h.nesterov) {
    m = h.momentum * m - lr * wd * w - lr * g;
    w += m;
  } else {
    g += wd * w;
    m = h.momentum * m + g;
    w -= lr * (g + h.momentum * m);
  }
}

struct AdamHyper {
  float lr, wd, beta1, beta2, eps, rescale, clip;
};

__device__ __forceinline__ void adam_step(float& w, float g, float& m, float& v, const AdamHyper& h, float lr,
                                          float wd) {
  g = clip_grad(g * h.rescale + wd * w, h.clip);
  m = h.beta1 * m + (1.f - h.beta1) * g;
  v = h.beta2 * v + (1.f - h.beta2) * g * g;
  w -= lr * m / (sqrtf(v) + h.eps);
}

struct LambHyper {
  float beta1, beta2, eps, rescale, clip;
  float bc1, bc2;  // 1-beta1^t, 1-beta2^t (1 when bias correction is off)
};

// phase 1: moments + raw update direction r (without lr / trust ratio)
__device__ __forceinline__ float lamb_dir(float w, float g, float& m, float& v, const LambHyper& h, float wd) {
  g = clip_grad(g * h.rescale, h.clip);
  m = h.beta1 * m + (1.f - h.beta1) * g;
  v = h.beta2 * v + (1.f - h.beta2) * g * g;
  const float mh = m / h.bc1, vh = v / h.bc2;
  return mh / (sqrtf(vh) + h.eps) + wd * w;
}

}  // namespace mxb
