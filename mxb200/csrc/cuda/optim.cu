// Fused optimizer update kernels (single tensor and multi-tensor "arena" forms).
//
// Reference kernels being replaced: SGDKernel / SGDMomKernel / MP_SGDKernel / MP_SGDMomKernel /
// AdamUpdate (src/operator/optimizer_op-inl.h:86-125,305-448,841-876), which launch one 256-thread
// grid-stride kernel per parameter tensor (and 3-4 launches for Adam).  Here one launch covers the whole
// parameter arena: tensors are laid out back to back (each padded to a 1024-element chunk) and a per-chunk
// table maps a chunk to its tensor's lr/wd multipliers, so ResNet-50's 161 tensors update in ONE launch,
// with 16-byte vector accesses, fp32 master weights + fp32 momentum and a low-precision (bf16) weight
// copy written in the same pass (and, optionally, the gradient zeroed in the same pass so no separate
// memset is needed before the next backward).
#include "common.cuh"
#include "optim_math.cuh"

namespace mxb {

constexpr int kChunk = 1024;   // arena granularity (elements)
constexpr int kOptThreads = 256;

template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&o)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&o)[4]) {
  float4 v = *reinterpret_cast<const float4*>(p);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&o)[4]) {
  uint2 v = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&v.x), b = *reinterpret_cast<__nv_bfloat162*>(&v.y);
  o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
}
template <>
__device__ __forceinline__ void load4<__half>(const __half* p, float (&o)[4]) {
  uint2 v = *reinterpret_cast<const uint2*>(p);
  __half2 a = *reinterpret_cast<__half2*>(&v.x), b = *reinterpret_cast<__half2*>(&v.y);
  o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
}
template <typename T>
__device__ __forceinline__ void store4(T* p, const float (&o)[4]);
template <>
__device__ __forceinline__ void store4<float>(float* p, const float (&o)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float (&o)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), b = __floats2bfloat162_rn(o[2], o[3]);
  uint2 v;
  v.x = *reinterpret_cast<uint32_t*>(&a);
  v.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = v;
}
template <>
__device__ __forceinline__ void store4<__half>(__half* p, const float (&o)[4]) {
  __half2 a = __floats2half2_rn(o[0], o[1]), b = __floats2half2_rn(o[2], o[3]);
  uint2 v;
  v.x = *reinterpret_cast<uint32_t*>(&a);
  v.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = v;
}

// ------------------------------------------------------------------------------------------------
// SGD / momentum / NAG.  WT = low-precision weight type, GT = gradient type, MT = momentum type.
// w32 (fp32 master) optional.  chunk table optional (null => uniform lr/wd, arbitrary n with scalar tail).
// ------------------------------------------------------------------------------------------------
template <typename WT, typename GT, typename MT>
__global__ void __launch_bounds__(kOptThreads)
sgd_kernel(WT* __restrict__ w, GT* __restrict__ g, MT* __restrict__ mom, float* __restrict__ w32, int64_t n,
           const int* __restrict__ chunk_tid, const float* __restrict__ lr_mult, const float* __restrict__ wd_mult,
           SgdHyper h, int zero_grad) {
  const int64_t nvec = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i << 2;
    float lr = h.lr, wd = h.wd;
    if (chunk_tid != nullptr) {
      const int t = __ldg(chunk_tid + (e / kChunk));
      if (t < 0) continue;  // padding chunk
      lr *= __ldg(lr_mult + t);
      wd *= __ldg(wd_mult + t);
    }
    float wv[4], gv[4], mv[4] = {0.f, 0.f, 0.f, 0.f};
    if (w32 != nullptr) load4<float>(w32 + e, wv);
    else load4<WT>(w + e, wv);
    load4<GT>(g + e, gv);
    if (h.has_mom) load4<MT>(mom + e, mv);
#pragma unroll
    for (int j = 0; j < 4; ++j) sgd_step(wv[j], gv[j], mv[j], h, lr, wd);
    if (w32 != nullptr) store4<float>(w32 + e, wv);
    if (w != nullptr) store4<WT>(w + e, wv);
    if (h.has_mom) store4<MT>(mom + e, mv);
    if (zero_grad) {
      const float z[4] = {0.f, 0.f, 0.f, 0.f};
      store4<GT>(g + e, z);
    }
  }
  // scalar tail (only without a chunk table; arenas are chunk padded)
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t e = (nvec << 2) + threadIdx.x;
    float wv = w32 ? w32[e] : to_f(w[e]);
    float mv = h.has_mom ? to_f(mom[e]) : 0.f;
    sgd_step(wv, to_f(g[e]), mv, h, h.lr, h.wd);
    if (w32) w32[e] = wv;
    if (w) w[e] = from_f<WT>(wv);
    if (h.has_mom) mom[e] = from_f<MT>(mv);
    if (zero_grad) g[e] = from_f<GT>(0.f);
  }
}

template <typename WT, typename GT>
__global__ void __launch_bounds__(kOptThreads)
adam_kernel(WT* __restrict__ w, GT* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            float* __restrict__ w32, int64_t n, const int* __restrict__ chunk_tid, const float* __restrict__ lr_mult,
            const float* __restrict__ wd_mult, AdamHyper h, int zero_grad) {
  const int64_t nvec = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i << 2;
    float lr = h.lr, wd = h.wd;
    if (chunk_tid != nullptr) {
      const int t = __ldg(chunk_tid + (e / kChunk));
      if (t < 0) continue;
      lr *= __ldg(lr_mult + t);
      wd *= __ldg(wd_mult + t);
    }
    float wv[4], gv[4], mv[4], vv[4];
    if (w32 != nullptr) load4<float>(w32 + e, wv);
    else load4<WT>(w + e, wv);
    load4<GT>(g + e, gv);
    load4<float>(m + e, mv);
    load4<float>(v + e, vv);
#pragma unroll
    for (int j = 0; j < 4; ++j) adam_step(wv[j], gv[j], mv[j], vv[j], h, lr, wd);
    if (w32 != nullptr) store4<float>(w32 + e, wv);
    if (w != nullptr) store4<WT>(w + e, wv);
    store4<float>(m + e, mv);
    store4<float>(v + e, vv);
    if (zero_grad) {
      const float z[4] = {0.f, 0.f, 0.f, 0.f};
      store4<GT>(g + e, z);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t e = (nvec << 2) + threadIdx.x;
    float wv = w32 ? w32[e] : to_f(w[e]);
    float mv = m[e], vv = v[e];
    adam_step(wv, to_f(g[e]), mv, vv, h, h.lr, h.wd);
    if (w32) w32[e] = wv;
    if (w) w[e] = from_f<WT>(wv);
    m[e] = mv; v[e] = vv;
    if (zero_grad) g[e] = from_f<GT>(0.f);
  }
}

// ------------------------------------------------------------------------------------------------
// LAMB, two phases.  Phase 1 updates m, v, writes the raw direction r (fp32) and accumulates per-tensor
// ||w||^2 and ||r||^2 (block reduce + one atomicAdd pair per block per tensor; a 1024-element chunk never
// straddles tensors, and a 256-thread block iteration covers exactly one chunk).
// Phase 2 applies w -= lr * trust(t) * r.
// ------------------------------------------------------------------------------------------------
template <typename WT, typename GT>
__global__ void __launch_bounds__(kOptThreads)
lamb_phase1_kernel(const WT* __restrict__ w, GT* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                   const float* __restrict__ w32, float* __restrict__ r_out, int64_t n_chunks,
                   const int* __restrict__ chunk_tid, const float* __restrict__ wd_mult, float* __restrict__ w_norm2,
                   float* __restrict__ r_norm2, LambHyper h, float wd_base, int zero_grad) {
  __shared__ float red[2][kOptThreads / 32];
  for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const int t = chunk_tid ? __ldg(chunk_tid + c) : 0;
    if (t < 0) continue;
    const float wd = wd_base * (wd_mult ? __ldg(wd_mult + t) : 1.f);
    const int64_t e = c * kChunk + threadIdx.x * 4;
    float wv[4], gv[4], mv[4], vv[4], rv[4];
    if (w32 != nullptr) load4<float>(w32 + e, wv);
    else load4<WT>(w + e, wv);
    load4<GT>(g + e, gv);
    load4<float>(m + e, mv);
    load4<float>(v + e, vv);
    float sw = 0.f, sr = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      rv[j] = lamb_dir(wv[j], gv[j], mv[j], vv[j], h, wd);
      sw += wv[j] * wv[j];
      sr += rv[j] * rv[j];
    }
    store4<float>(m + e, mv);
    store4<float>(v + e, vv);
    store4<float>(r_out + e, rv);
    if (zero_grad) {
      const float z[4] = {0.f, 0.f, 0.f, 0.f};
      store4<GT>(g + e, z);
    }
    sw = warp_sum(sw);
    sr = warp_sum(sr);
    if ((threadIdx.x & 31) == 0) {
      red[0][threadIdx.x >> 5] = sw;
      red[1][threadIdx.x >> 5] = sr;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      float a = threadIdx.x < kOptThreads / 32 ? red[0][threadIdx.x] : 0.f;
      float b = threadIdx.x < kOptThreads / 32 ? red[1][threadIdx.x] : 0.f;
      a = warp_sum(a);
      b = warp_sum(b);
      if (threadIdx.x == 0) {
        atomicAdd(w_norm2 + t, a);
        atomicAdd(r_norm2 + t, b);
      }
    }
    __syncthreads();
  }
}

template <typename WT>
__global__ void __launch_bounds__(kOptThreads)
lamb_phase2_kernel(WT* __restrict__ w, float* __restrict__ w32, const float* __restrict__ r, int64_t n_chunks,
                   const int* __restrict__ chunk_tid, const float* __restrict__ lr_mult,
                   const float* __restrict__ w_norm2, const float* __restrict__ r_norm2, float lr_base,
                   float lower_bound, float upper_bound) {
  for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const int t = chunk_tid ? __ldg(chunk_tid + c) : 0;
    if (t < 0) continue;
    float wn = sqrtf(__ldg(w_norm2 + t)), rn = sqrtf(__ldg(r_norm2 + t));
    if (lower_bound >= 0.f) wn = fmaxf(wn, lower_bound);
    if (upper_bound >= 0.f) wn = fminf(wn, upper_bound);
    const float ratio = (wn > 0.f && rn > 0.f) ? wn / rn : 1.f;
    const float lr = lr_base * (lr_mult ? __ldg(lr_mult + t) : 1.f) * ratio;
    const int64_t e = c * kChunk + threadIdx.x * 4;
    float wv[4], rv[4];
    if (w32 != nullptr) load4<float>(w32 + e, wv);
    else load4<WT>(w + e, wv);
    load4<float>(r + e, rv);
#pragma unroll
    for (int j = 0; j < 4; ++j) wv[j] -= lr * rv[j];
    if (w32 != nullptr) store4<float>(w32 + e, wv);
    if (w != nullptr) store4<WT>(w + e, wv);
  }
}

// out = sum_i in[i]  (ElementwiseSum, src/ndarray/ndarray_function-inl.h:388-430), up to 8 inputs, fp32 accumulate
struct SumArgs {
  const void* in[8];
  int n_in;
};
template <typename T>
__global__ void __launch_bounds__(kOptThreads)
elementwise_sum_kernel(T* __restrict__ out, SumArgs a, int64_t n, float scale) {
  const int64_t nvec = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < a.n_in; ++k) {
      float v[4];
      load4<T>(reinterpret_cast<const T*>(a.in[k]) + (i << 2), v);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] *= scale;
    store4<T>(out + (i << 2), acc);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t e = (nvec << 2) + threadIdx.x;
    float acc = 0.f;
    for (int k = 0; k < a.n_in; ++k) acc += to_f(reinterpret_cast<const T*>(a.in[k])[e]);
    out[e] = from_f<T>(acc * scale);
  }
}

template <typename F>
static void dispatch_dtype(int dt, F&& f) {
  switch (dt) {
    case kF32: f(float{}); break;
    case kBF16: f(__nv_bfloat16{}); break;
    case kF16: f(__half{}); break;
    default: throw std::runtime_error("unsupported dtype code " + std::to_string(dt));
  }
}

static void check_align(const void* p, const char* what) {
  if (p && (reinterpret_cast<uintptr_t>(p) & 15)) throw std::runtime_error(std::string(what) + " not 16B aligned");
}

}  // namespace mxb

using namespace mxb;

MXB_API int MXBSgdUpdate(void* w, int w_dt, void* g, int g_dt, void* mom, int mom_dt, float* w32, int64_t n,
                         const int* chunk_tid, const float* lr_mult, const float* wd_mult, float lr, float wd,
                         float momentum, float rescale, float clip, int nesterov, int zero_grad, void* stream) {
  MXB_TRY
  check_align(w, "weight"); check_align(g, "grad"); check_align(mom, "mom"); check_align(w32, "weight32");
  SgdHyper h{lr, wd, momentum, rescale, clip, nesterov, mom != nullptr};
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = grid_for(n, kOptThreads, 4);
  dispatch_dtype(w_dt, [&](auto wt) {
    using WT = decltype(wt);
    dispatch_dtype(g_dt, [&](auto gt) {
      using GT = decltype(gt);
      if (mom == nullptr || mom_dt == kF32) {
        sgd_kernel<WT, GT, float><<<grid, kOptThreads, 0, s>>>((WT*)w, (GT*)g, (float*)mom, w32, n, chunk_tid, lr_mult,
                                                              wd_mult, h, zero_grad);
      } else {
        if (mom_dt != w_dt) throw std::runtime_error("momentum dtype must be fp32 or the weight dtype");
        sgd_kernel<WT, GT, WT><<<grid, kOptThreads, 0, s>>>((WT*)w, (GT*)g, (WT*)mom, w32, n, chunk_tid, lr_mult,
                                                           wd_mult, h, zero_grad);
      }
    });
  });
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}

MXB_API int MXBAdamUpdate(void* w, int w_dt, void* g, int g_dt, float* m, float* v, float* w32, int64_t n,
                          const int* chunk_tid, const float* lr_mult, const float* wd_mult, float lr, float wd,
                          float beta1, float beta2, float eps, float rescale, float clip, int zero_grad, void* stream) {
  MXB_TRY
  check_align(w, "weight"); check_align(g, "grad"); check_align(m, "mean"); check_align(v, "var");
  AdamHyper h{lr, wd, beta1, beta2, eps, rescale, clip};
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = grid_for(n, kOptThreads, 4);
  dispatch_dtype(w_dt, [&](auto wt) {
    using WT = decltype(wt);
    dispatch_dtype(g_dt, [&](auto gt) {
      using GT = decltype(gt);
      adam_kernel<WT, GT><<<grid, kOptThreads, 0, s>>>((WT*)w, (GT*)g, m, v, w32, n, chunk_tid, lr_mult, wd_mult, h,
                                                      zero_grad);
    });
  });
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}

// n must be a multiple of 1024 (arena) -- single tensors are handled by padding on the python side.
MXB_API int MXBLambPhase1(const void* w, int w_dt, void* g, int g_dt, float* m, float* v, const float* w32, float* r,
                          int64_t n, const int* chunk_tid, const float* wd_mult, float* w_norm2, float* r_norm2,
                          float beta1, float beta2, float eps, float rescale, float clip, float bc1, float bc2,
                          float wd, int zero_grad, void* stream) {
  MXB_TRY
  MXB_CHECK(n % kChunk == 0, "LAMB arena size must be a multiple of 1024");
  LambHyper h{beta1, beta2, eps, rescale, clip, bc1, bc2};
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t nc = n / kChunk;
  const int grid = (int)std::min<int64_t>(nc, (int64_t)kNumSMs * 8);
  dispatch_dtype(w_dt, [&](auto wt) {
    using WT = decltype(wt);
    dispatch_dtype(g_dt, [&](auto gt) {
      using GT = decltype(gt);
      lamb_phase1_kernel<WT, GT><<<grid, kOptThreads, 0, s>>>((const WT*)w, (GT*)g, m, v, w32, r, nc, chunk_tid,
                                                             wd_mult, w_norm2, r_norm2, h, wd, zero_grad);
    });
  });
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}

MXB_API int MXBLambPhase2(void* w, int w_dt, float* w32, const float* r, int64_t n, const int* chunk_tid,
                          const float* lr_mult, const float* w_norm2, const float* r_norm2, float lr,
                          float lower_bound, float upper_bound, void* stream) {
  MXB_TRY
  MXB_CHECK(n % kChunk == 0, "LAMB arena size must be a multiple of 1024");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t nc = n / kChunk;
  const int grid = (int)std::min<int64_t>(nc, (int64_t)kNumSMs * 8);
  dispatch_dtype(w_dt, [&](auto wt) {
    using WT = decltype(wt);
    lamb_phase2_kernel<WT><<<grid, kOptThreads, 0, s>>>((WT*)w, w32, r, nc, chunk_tid, lr_mult, w_norm2, r_norm2, lr,
                                                       lower_bound, upper_bound);
  });
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}

MXB_API int MXBElementwiseSum(void* out, int dt, const void* const* ins, int n_in, int64_t n, float scale,
                              void* stream) {
  MXB_TRY
  MXB_CHECK(n_in >= 1 && n_in <= 8, "ElementwiseSum supports 1..8 inputs per launch");
  SumArgs a;
  a.n_in = n_in;
  for (int i = 0; i < n_in; ++i) {
    check_align(ins[i], "input");
    a.in[i] = ins[i];
  }
  check_align(out, "output");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = grid_for(n, kOptThreads, 4);
  dispatch_dtype(dt, [&](auto t) {
    using T = decltype(t);
    elementwise_sum_kernel<T><<<grid, kOptThreads, 0, s>>>((T*)out, a, n, scale);
  });
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}
