// Input pipeline tail on the device: uint8 NHWC image batch -> normalised bf16 (or fp32) tensor in one pass.
//
// Reference: the ImageRecordIter pipeline normalises on the CPU (src/io/iter_normalize.h: (x - mean) / std per channel)
// and the fp16 examples then Cast on the GPU (example/image-classification/symbols/resnet.py: mx.sym.Cast) -- two
// more full passes over the batch.  Here the raw uint8 batch crosses PCIe/C2C (4x fewer bytes than fp32) and one
// kernel reads 16 pixels-channels (16 B) and writes 32 B of bf16 per thread.
#include "common.cuh"

namespace mxb {
namespace image {

constexpr int kThreads = 256;

template <typename OutT>
__global__ void __launch_bounds__(kThreads)
normalize_u8_kernel(const uint8_t* __restrict__ src, OutT* __restrict__ dst, int64_t n, int C, const float* __restrict__ mean,
                    const float* __restrict__ istd) {
  __shared__ float sm[64], ss[64];
  if (threadIdx.x < C) {
    sm[threadIdx.x] = mean[threadIdx.x];
    ss[threadIdx.x] = istd[threadIdx.x];
  }
  __syncthreads();
  const int64_t nvec = n / 16;
  for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * kThreads) {
    const uint4 u = ldg16_stream(src + v * 16);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    int c = (int)((v * 16) % C);
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float x = (float)((w[i >> 2] >> (8 * (i & 3))) & 0xffu);
      f[i] = (x - sm[c]) * ss[c];
      if (++c == C) c = 0;
    }
    if constexpr (sizeof(OutT) == 2) {
      uint4 o[2];
      __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      stg16(dst + v * 16, o[0]);
      stg16(dst + v * 16 + 8, o[1]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(dst + v * 16 + 4 * i) = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
    }
  }
  // tail (n not a multiple of 16)
  for (int64_t i = nvec * 16 + (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const int c = (int)(i % C);
    dst[i] = from_f<OutT>(((float)src[i] - sm[c]) * ss[c]);
  }
}

}  // namespace image
}  // namespace mxb

using namespace mxb;

// src: uint8 [n] with the channel as the fastest dimension (NHWC), C <= 64; dst: bf16 / fp32 [n]; mean, istd: fp32 [C]
MXB_API int MXBNormalizeU8(const void* src, void* dst, int out_dt, int64_t n, int C, const float* mean, const float* istd,
                           void* stream) {
  MXB_TRY
  MXB_CHECK(C >= 1 && C <= 64, "normalize: 1..64 channels");
  MXB_CHECK((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "normalize: 16B alignment");
  if (n == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = grid_for(n / 16 + 1, image::kThreads, 2, 16);
  if (out_dt == kBF16)
    image::normalize_u8_kernel<__nv_bfloat16><<<grid, image::kThreads, 0, s>>>((const uint8_t*)src, (__nv_bfloat16*)dst, n, C, mean, istd);
  else if (out_dt == kF32)
    image::normalize_u8_kernel<float><<<grid, image::kThreads, 0, s>>>((const uint8_t*)src, (float*)dst, n, C, mean, istd);
  else
    throw std::runtime_error("normalize: output must be bf16 or fp32");
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}
