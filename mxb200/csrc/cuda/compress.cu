// 2-bit gradient compression with error-feedback residual (device side of kvstore.set_gradient_compression).
//
// Reference semantics: src/kvstore/gradient_compression-inl.h:40-127 (quantize_2bit / dequantize_2bit, one thread
// per 16 values, launched through the generic Kernel<>::Launch with 256-thread blocks).  Wire format (bit-exact
// with the reference so compressed pushes interoperate with its NumPy oracle, tests/nightly/test_kvstore.py:33-77):
// 16 values -> one 32-bit word; value i lives in byte (i >> 2) (byte 0 = least significant), bit pair (i & 3)
// counted from the byte's MSB; 11 = +threshold, 10 = -threshold, 00 = 0.
//
// B200 version: one thread still owns one output word (the word is the atomic unit of the format) but moves its
// 16 inputs with four 16-byte loads/stores (64 B of gradient + 64 B of residual read, 64 B residual + 4 B code
// written per thread), so a warp touches 2 KB contiguous per array and the kernel streams at HBM rate.
#include "common.cuh"

namespace mxb {
namespace compress {

constexpr int kThreads = 256;

template <typename T>
__device__ __forceinline__ void load16(const T* p, int64_t base, int64_t n, float* v);

template <>
__device__ __forceinline__ void load16<float>(const float* p, int64_t base, int64_t n, float* v) {
  if (base + 16 <= n && (reinterpret_cast<uintptr_t>(p + base) & 15) == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 f = *reinterpret_cast<const float4*>(p + base + 4 * q);
      v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (base + i < n) ? p[base + i] : 0.f;
  }
}
template <>
__device__ __forceinline__ void load16<__nv_bfloat16>(const __nv_bfloat16* p, int64_t base, int64_t n, float* v) {
  if (base + 16 <= n && (reinterpret_cast<uintptr_t>(p + base) & 15) == 0) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const uint4 u = *reinterpret_cast<const uint4*>(p + base + 8 * q);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[8 * q + 2 * k] = __low2float(h[k]);
        v[8 * q + 2 * k + 1] = __high2float(h[k]);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (base + i < n) ? __bfloat162float(p[base + i]) : 0.f;
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
quantize_2bit_kernel(const T* __restrict__ grad, float* __restrict__ residual, uint32_t* __restrict__ out, int64_t n,
                     int64_t words, float thr) {
  for (int64_t w = (int64_t)blockIdx.x * kThreads + threadIdx.x; w < words; w += (int64_t)gridDim.x * kThreads) {
    const int64_t base = w * 16;
    float g[16], r[16];
    load16<T>(grad, base, n, g);
    load16<float>(residual, base, n, r);
    uint32_t word = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float x = r[i] + g[i];
      uint32_t code = 0;
      if (x >= thr) {
        code = 3u;
        x -= thr;
      } else if (x <= -thr) {
        code = 2u;
        x += thr;
      }
      r[i] = x;
      word |= code << (8 * (i >> 2) + 6 - 2 * (i & 3));
    }
    out[w] = word;
    if (base + 16 <= n && (reinterpret_cast<uintptr_t>(residual + base) & 15) == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(residual + base + 4 * q) = make_float4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (base + i < n) residual[base + i] = r[i];
    }
  }
}

template <bool ACC>
__global__ void __launch_bounds__(kThreads)
dequantize_2bit_kernel(const uint32_t* __restrict__ in, float* __restrict__ out, int64_t n, int64_t words, float thr) {
  for (int64_t w = (int64_t)blockIdx.x * kThreads + threadIdx.x; w < words; w += (int64_t)gridDim.x * kThreads) {
    const uint32_t word = in[w];
    const int64_t base = w * 16;
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const uint32_t code = (word >> (8 * (i >> 2) + 6 - 2 * (i & 3))) & 3u;
      v[i] = code == 3u ? thr : (code == 2u ? -thr : 0.f);
    }
    if (base + 16 <= n && (reinterpret_cast<uintptr_t>(out + base) & 15) == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 f = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        if (ACC) {
          const float4 o = *reinterpret_cast<const float4*>(out + base + 4 * q);
          f.x += o.x; f.y += o.y; f.z += o.z; f.w += o.w;
        }
        *reinterpret_cast<float4*>(out + base + 4 * q) = f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (base + i < n) out[base + i] = ACC ? out[base + i] + v[i] : v[i];
    }
  }
}

}  // namespace compress
}  // namespace mxb

using namespace mxb;

// grad: fp32 or bf16 [n]; residual: fp32 [n] (updated in place); out: int32 words [ceil(n/16)]
MXB_API int MXBQuantize2Bit(const void* grad, int dt, float* residual, void* out, int64_t n, float threshold,
                            void* stream) {
  MXB_TRY
  MXB_CHECK(threshold > 0.f, "threshold must be greater than 0 for two bit compression");
  const int64_t words = (n + 15) / 16;
  if (words == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = grid_for(words, compress::kThreads, 1, 16);
  if (dt == kF32)
    compress::quantize_2bit_kernel<float><<<grid, compress::kThreads, 0, s>>>((const float*)grad, residual, (uint32_t*)out, n, words, threshold);
  else if (dt == kBF16)
    compress::quantize_2bit_kernel<__nv_bfloat16><<<grid, compress::kThreads, 0, s>>>((const __nv_bfloat16*)grad, residual, (uint32_t*)out, n, words, threshold);
  else
    throw std::runtime_error("quantize_2bit: gradient must be fp32 or bf16");
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}

// words [ceil(n/16)] -> out fp32 [n]; accumulate != 0 adds into out (server-side merge of several workers' pushes)
MXB_API int MXBDequantize2Bit(const void* words_in, float* out, int64_t n, float threshold, int accumulate,
                              void* stream) {
  MXB_TRY
  const int64_t words = (n + 15) / 16;
  if (words == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = grid_for(words, compress::kThreads, 1, 16);
  if (accumulate)
    compress::dequantize_2bit_kernel<true><<<grid, compress::kThreads, 0, s>>>((const uint32_t*)words_in, out, n, words, threshold);
  else
    compress::dequantize_2bit_kernel<false><<<grid, compress::kThreads, 0, s>>>((const uint32_t*)words_in, out, n, words, threshold);
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}
