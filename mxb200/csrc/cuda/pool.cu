// Max pooling for channels-last activations (forward with argmax byte-map, backward as a gather).
//
// Reference op: Pooling (src/operator/nn/pool.cuh:96-650, one thread per output element, NCHW) / cuDNN pooling.
// The stem's 3x3/s2 max-pool runs on the largest activation of the network (N x 112 x 112 x 64); the stock
// channels-last kernels take >2 ms fwd+bwd at batch 256, ~10x the memory roofline.  Here each thread owns
// 8 channels (one 16-byte vector): the forward keeps a 1-byte argmax (window position 0..k*k-1) per output
// element, and the backward is a *gather* over the <= ceil(k/s)^2 windows that can contain an input pixel,
// so there are no atomics and every dx element is written exactly once.
#include "common.cuh"

namespace mxb {
namespace pool {

constexpr int kThreads = 256;

template <typename T>
struct V {
  static constexpr int N = 16 / sizeof(T);
};
template <typename T>
__device__ __forceinline__ void unpack(const uint4& u, float* o);
template <>
__device__ __forceinline__ void unpack<__nv_bfloat16>(const uint4& u, float* o) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[2 * i] = __low2float(p[i]);
    o[2 * i + 1] = __high2float(p[i]);
  }
}
template <>
__device__ __forceinline__ void unpack<float>(const uint4& u, float* o) {
  o[0] = __uint_as_float(u.x); o[1] = __uint_as_float(u.y); o[2] = __uint_as_float(u.z); o[3] = __uint_as_float(u.w);
}
template <typename T>
__device__ __forceinline__ uint4 pack(const float* o);
template <>
__device__ __forceinline__ uint4 pack<__nv_bfloat16>(const float* o) {
  uint4 u;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
  return u;
}
template <>
__device__ __forceinline__ uint4 pack<float>(const float* o) {
  return make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]));
}

struct Geo {
  int N, H, W, C, HO, WO, k, s, p;
};

// KS > 0: compile-time window (KS x KS, stride KS-1 ... the ResNet stem uses 3x3 / stride 2): every load of a window is
// issued before the first comparison (memory-level parallelism), out-of-image taps read as -inf.
template <typename T, int K, int S>
__global__ void __launch_bounds__(kThreads, 2)
maxpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, uint8_t* __restrict__ idx, Geo g) {
  constexpr int VN = V<T>::N;
  const int k = K > 0 ? K : g.k, s = S > 0 ? S : g.s;
  const int CV = g.C / VN;
  const int64_t total = (int64_t)g.N * g.HO * g.WO * CV;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int cv = (int)(i % CV);
    int64_t r = i / CV;
    const int wo = (int)(r % g.WO); r /= g.WO;
    const int ho = (int)(r % g.HO);
    const int n = (int)(r / g.HO);
    float best[VN];
    int arg[VN];
#pragma unroll
    for (int j = 0; j < VN; ++j) { best[j] = -INFINITY; arg[j] = 0; }
    const T* base = x + ((int64_t)n * g.H * g.W) * g.C + (int64_t)cv * VN;
    if constexpr (K > 0) {
      uint4 u[K * K];
      bool ok[K * K];
#pragma unroll
      for (int kh = 0; kh < K; ++kh)
#pragma unroll
        for (int kw = 0; kw < K; ++kw) {
          const int h = ho * s - g.p + kh, w = wo * s - g.p + kw;
          ok[kh * K + kw] = (unsigned)h < (unsigned)g.H && (unsigned)w < (unsigned)g.W;
          u[kh * K + kw] = ok[kh * K + kw] ? ldg16(base + ((int64_t)h * g.W + w) * g.C) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
      for (int t = 0; t < K * K; ++t) {
        if (!ok[t]) continue;
        float v[VN];
        unpack<T>(u[t], v);
#pragma unroll
        for (int j = 0; j < VN; ++j)
          if (v[j] > best[j]) { best[j] = v[j]; arg[j] = t; }
      }
    } else {
      for (int kh = 0; kh < k; ++kh) {
        const int h = ho * s - g.p + kh;
        if (h < 0 || h >= g.H) continue;
        for (int kw = 0; kw < k; ++kw) {
          const int w = wo * s - g.p + kw;
          if (w < 0 || w >= g.W) continue;
          float v[VN];
          unpack<T>(ldg16(base + ((int64_t)h * g.W + w) * g.C), v);
#pragma unroll
          for (int j = 0; j < VN; ++j)
            if (v[j] > best[j]) { best[j] = v[j]; arg[j] = kh * k + kw; }
        }
      }
    }
    stg16(y + i * VN, pack<T>(best));
    // VN argmax bytes, contiguous per output vector
    if (VN == 8) {
      uint2 a;
      a.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
      a.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
      *reinterpret_cast<uint2*>(idx + i * 8) = a;
    } else {
      uint32_t a = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
      *reinterpret_cast<uint32_t*>(idx + i * 4) = a;
    }
  }
}

template <typename T, int K, int S>
__global__ void __launch_bounds__(kThreads, 2)
maxpool_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ idx, T* __restrict__ dx, Geo g) {
  constexpr int VN = V<T>::N;
  constexpr int kMaxWin = K > 0 ? (K + S - 1) / S : 1;     // candidate windows per axis that can cover one input pixel
  const int k = K > 0 ? K : g.k, s = S > 0 ? S : g.s;
  const int CV = g.C / VN;
  const int64_t total = (int64_t)g.N * g.H * g.W * CV;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int cv = (int)(i % CV);
    int64_t r = i / CV;
    const int w = (int)(r % g.W); r /= g.W;
    const int h = (int)(r % g.H);
    const int n = (int)(r / g.H);
    float acc[VN];
#pragma unroll
    for (int j = 0; j < VN; ++j) acc[j] = 0.f;
    // outputs whose window covers (h, w): ho in [ceil((h+p-k+1)/s), floor((h+p)/s)]
    const int ho_hi = min(g.HO - 1, (h + g.p) / s);
    const int ho_lo = max(0, (h + g.p - k + s) / s);
    const int wo_hi = min(g.WO - 1, (w + g.p) / s);
    const int wo_lo = max(0, (w + g.p - k + s) / s);
    auto fold = [&](const uint4& uv, const uint2& a, uint32_t a4, int code) {
      float v[VN];
      unpack<T>(uv, v);
      if (VN == 8) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (((a.x >> (8 * j)) & 0xff) == (uint32_t)code) acc[j] += v[j];
          if (((a.y >> (8 * j)) & 0xff) == (uint32_t)code) acc[4 + j] += v[4 + j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (((a4 >> (8 * j)) & 0xff) == (uint32_t)code) acc[j] += v[j];
      }
    };
    if constexpr (K > 0) {
      uint4 uv[kMaxWin * kMaxWin];
      uint2 ua[kMaxWin * kMaxWin];
      uint32_t u4[kMaxWin * kMaxWin];
      int code[kMaxWin * kMaxWin];
#pragma unroll
      for (int a = 0; a < kMaxWin; ++a)
#pragma unroll
        for (int b = 0; b < kMaxWin; ++b) {
          const int ho = ho_lo + a, wo = wo_lo + b, t = a * kMaxWin + b;
          const bool ok = ho <= ho_hi && wo <= wo_hi;
          code[t] = ok ? (h + g.p - ho * s) * k + (w + g.p - wo * s) : -1;
          const int64_t o = (((int64_t)n * g.HO + (ok ? ho : 0)) * g.WO + (ok ? wo : 0)) * CV + cv;
          uv[t] = ok ? ldg16(dy + o * VN) : make_uint4(0, 0, 0, 0);
          ua[t] = make_uint2(0, 0);
          u4[t] = 0;
          if (ok) {
            if (VN == 8) ua[t] = *reinterpret_cast<const uint2*>(idx + o * 8);
            else u4[t] = *reinterpret_cast<const uint32_t*>(idx + o * 4);
          }
        }
#pragma unroll
      for (int t = 0; t < kMaxWin * kMaxWin; ++t)
        if (code[t] >= 0) fold(uv[t], ua[t], u4[t], code[t]);
    } else {
      for (int ho = ho_lo; ho <= ho_hi; ++ho) {
        const int kh = h + g.p - ho * s;
        for (int wo = wo_lo; wo <= wo_hi; ++wo) {
          const int kw = w + g.p - wo * s;
          const int64_t o = (((int64_t)n * g.HO + ho) * g.WO + wo) * CV + cv;
          uint2 a = make_uint2(0, 0);
          uint32_t a4 = 0;
          if (VN == 8) a = *reinterpret_cast<const uint2*>(idx + o * 8);
          else a4 = *reinterpret_cast<const uint32_t*>(idx + o * 4);
          fold(ldg16(dy + o * VN), a, a4, kh * k + kw);
        }
      }
    }
    stg16(dx + i * VN, pack<T>(acc));
  }
}

}  // namespace pool
}  // namespace mxb

using namespace mxb;
using namespace mxb::pool;

MXB_API int MXBMaxPoolNhwcFwd(const void* x, void* y, void* idx, int dt, int N, int H, int W, int C, int HO, int WO,
                              int k, int s, int p, void* stream) {
  MXB_TRY
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Geo g{N, H, W, C, HO, WO, k, s, p};
  MXB_CHECK(k * k <= 255, "kernel too large for the byte argmax map");
  if (dt == kBF16) {
    MXB_CHECK(C % 8 == 0, "maxpool: C must be a multiple of 8 for bf16");
    const int64_t total = (int64_t)N * HO * WO * (C / 8);
    if (k == 3 && s == 2)
      maxpool_fwd_kernel<__nv_bfloat16, 3, 2><<<grid_for(total, kThreads, 1, 16), kThreads, 0, st>>>(
          (const __nv_bfloat16*)x, (__nv_bfloat16*)y, (uint8_t*)idx, g);
    else
      maxpool_fwd_kernel<__nv_bfloat16, 0, 0><<<grid_for(total, kThreads, 1, 16), kThreads, 0, st>>>(
          (const __nv_bfloat16*)x, (__nv_bfloat16*)y, (uint8_t*)idx, g);
  } else if (dt == kF32) {
    MXB_CHECK(C % 4 == 0, "maxpool: C must be a multiple of 4 for fp32");
    const int64_t total = (int64_t)N * HO * WO * (C / 4);
    if (k == 3 && s == 2)
      maxpool_fwd_kernel<float, 3, 2><<<grid_for(total, kThreads, 1, 16), kThreads, 0, st>>>((const float*)x, (float*)y, (uint8_t*)idx, g);
    else
      maxpool_fwd_kernel<float, 0, 0><<<grid_for(total, kThreads, 1, 16), kThreads, 0, st>>>((const float*)x, (float*)y, (uint8_t*)idx, g);
  } else {
    throw std::runtime_error("maxpool: dtype must be bf16 or fp32");
  }
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}

MXB_API int MXBMaxPoolNhwcBwd(const void* dy, const void* idx, void* dx, int dt, int N, int H, int W, int C, int HO,
                              int WO, int k, int s, int p, void* stream) {
  MXB_TRY
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Geo g{N, H, W, C, HO, WO, k, s, p};
  if (dt == kBF16) {
    const int64_t total = (int64_t)N * H * W * (C / 8);
    if (k == 3 && s == 2)
      maxpool_bwd_kernel<__nv_bfloat16, 3, 2><<<grid_for(total, kThreads, 1, 16), kThreads, 0, st>>>(
          (const __nv_bfloat16*)dy, (const uint8_t*)idx, (__nv_bfloat16*)dx, g);
    else
      maxpool_bwd_kernel<__nv_bfloat16, 0, 0><<<grid_for(total, kThreads, 1, 16), kThreads, 0, st>>>(
          (const __nv_bfloat16*)dy, (const uint8_t*)idx, (__nv_bfloat16*)dx, g);
  } else if (dt == kF32) {
    const int64_t total = (int64_t)N * H * W * (C / 4);
    if (k == 3 && s == 2)
      maxpool_bwd_kernel<float, 3, 2><<<grid_for(total, kThreads, 1, 16), kThreads, 0, st>>>((const float*)dy, (const uint8_t*)idx, (float*)dx, g);
    else
      maxpool_bwd_kernel<float, 0, 0><<<grid_for(total, kThreads, 1, 16), kThreads, 0, st>>>((const float*)dy, (const uint8_t*)idx, (float*)dx, g);
  } else {
    throw std::runtime_error("maxpool: dtype must be bf16 or fp32");
  }
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}

// ------------------------------------------------------------------ spatial sub-sampling x[:, ::s, ::s, :] (NHWC)
// The strided 1x1 "downsample" convolutions of ResNet read every s-th pixel.  Forward = gather into a dense tensor (the
// GEMM's A operand); backward writes the FULL input gradient in one pass (dy at the sampled pixels, zero elsewhere)
// instead of a zero-fill kernel plus a strided scatter.  16-byte vectors, byte-granular (dtype agnostic).
namespace mxb {
namespace pool {

__global__ void __launch_bounds__(kThreads)
subsample_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int CV, int HO, int WO, int s) {
  const int64_t total = (int64_t)N * HO * WO * CV;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int c = (int)(i % CV);
    int64_t r = i / CV;
    const int wo = (int)(r % WO); r /= WO;
    const int ho = (int)(r % HO);
    const int n = (int)(r / HO);
    y[i] = ldg16_stream(x + (((int64_t)n * H + (int64_t)ho * s) * W + (int64_t)wo * s) * CV + c);
  }
}

__global__ void __launch_bounds__(kThreads)
subsample_bwd_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dx, int N, int H, int W, int CV, int HO, int WO, int s) {
  const int64_t total = (int64_t)N * H * W * CV;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int c = (int)(i % CV);
    int64_t r = i / CV;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (h % s == 0 && w % s == 0 && h / s < HO && w / s < WO)
      v = ldg16_stream(dy + (((int64_t)n * HO + h / s) * WO + w / s) * CV + c);
    dx[i] = v;
  }
}

}  // namespace pool
}  // namespace mxb

// x: [N,H,W,C] contiguous with C*elem_bytes a multiple of 16; y: [N,HO,WO,C], HO = ceil(H/s), WO = ceil(W/s)
MXB_API int MXBSubsampleNhwc(const void* src, void* dst, int elem_bytes, int N, int H, int W, int C, int s, int backward,
                             void* stream) {
  MXB_TRY
  MXB_CHECK((C * elem_bytes) % 16 == 0, "subsample: C * sizeof(T) must be a multiple of 16");
  MXB_CHECK(s >= 1, "subsample: stride >= 1");
  const int CV = C * elem_bytes / 16;
  const int HO = (H + s - 1) / s, WO = (W + s - 1) / s;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!backward) {
    const int64_t total = (int64_t)N * HO * WO * CV;
    if (total == 0) return 0;
    pool::subsample_fwd_kernel<<<grid_for(total, pool::kThreads, 4, 16), pool::kThreads, 0, st>>>(
        (const uint4*)src, (uint4*)dst, N, H, W, CV, HO, WO, s);
  } else {
    const int64_t total = (int64_t)N * H * W * CV;
    if (total == 0) return 0;
    pool::subsample_bwd_kernel<<<grid_for(total, pool::kThreads, 4, 16), pool::kThreads, 0, st>>>(
        (const uint4*)src, (uint4*)dst, N, H, W, CV, HO, WO, s);
  }
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}
