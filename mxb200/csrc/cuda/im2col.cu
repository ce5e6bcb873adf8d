// Patch extraction (im2col) for small-channel stem convolutions on channels-last bf16 images.
//
// Why: the ResNet stem (7x7 / stride 2 / 3 input channels; reference: Convolution op -> cuDNN,
// src/operator/nn/convolution.cu, falling back to im2col + GEMM, nn/im2col.cuh:79-180) has too few input channels
// for the tensor-core implicit-GEMM kernels of the library: measured on B200 at batch 256 the library takes 1.48 ms
// forward + 0.77 ms wgrad, 10-20x off the memory roofline of that layer.  Here the stem becomes
//     A[M = N*OH*OW, Kp] = patches (k = (kh*KW + kw)*C + c, zero padded to Kp = round_up(KH*KW*C, 8))
//     Y = A . Wp^T      (our tcgen05 GEMM, BatchNorm statistics in its epilogue)
//     dW = dY^T . A     (same GEMM, TN mode, fp32 split-K reduce straight into the gradient arena)
// A is written once (16-byte stores, fully coalesced: consecutive threads own consecutive 16-byte chunks of a row)
// and kept for the backward pass; the input image is read through L1/L2 (each pixel is used by ~12 patches).
#include "common.cuh"

namespace mxb {
namespace im2col {

constexpr int kThreads = 256;
constexpr int kMaxK = 512;

struct Div {   // n / d via multiply-high (d is launch-constant)
  uint32_t d, mul, shift;
  __device__ __forceinline__ uint32_t div(uint32_t n) const {
    return (uint32_t)(((uint64_t)__umulhi(n, mul) + n) >> shift);
  }
};

static Div make_div(uint32_t d) {
  Div f;
  f.d = d;
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;
  f.shift = s;
  f.mul = (uint32_t)((((1ull << 32) * ((1ull << s) - d)) / d) + 1);
  return f;
}

struct Params {
  const __nv_bfloat16* x;   // [N, H, W, C]
  __nv_bfloat16* out;       // [M, Kp]
  int H, W, C, KH, KW, stride, pad, OH, OW, K, Kp, chunks;
  int64_t total;            // M * chunks
  Div d_chunks, d_ohow, d_ow;
};

__global__ void __launch_bounds__(kThreads) im2col_kernel(const Params p) {
  // per-k lookup table: dh | dw << 8 | c << 16 | valid << 24
  __shared__ uint32_t tab[kMaxK];
  for (int k = threadIdx.x; k < p.Kp; k += kThreads) {
    uint32_t e = 0;
    if (k < p.K) {
      const int kh = k / (p.KW * p.C), r = k - kh * (p.KW * p.C);
      const int kw = r / p.C, c = r - kw * p.C;
      e = (uint32_t)kh | ((uint32_t)kw << 8) | ((uint32_t)c << 16) | (1u << 24);
    }
    tab[k] = e;
  }
  __syncthreads();
  const unsigned short* xs = reinterpret_cast<const unsigned short*>(p.x);
  for (int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x; idx < p.total; idx += (int64_t)gridDim.x * kThreads) {
    // idx < 2^32 is checked on the host
    const uint32_t m = p.d_chunks.div((uint32_t)idx);
    const uint32_t j = (uint32_t)idx - m * p.d_chunks.d;
    const uint32_t n = p.d_ohow.div(m);
    const uint32_t rem = m - n * p.d_ohow.d;
    const uint32_t oh = p.d_ow.div(rem);
    const uint32_t ow = rem - oh * p.d_ow.d;
    const int ih0 = (int)oh * p.stride - p.pad, iw0 = (int)ow * p.stride - p.pad;
    const unsigned short* img = xs + (int64_t)n * p.H * p.W * p.C;
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const uint32_t t = tab[j * 8 + e];
      const int ih = ih0 + (int)(t & 0xff), iw = iw0 + (int)((t >> 8) & 0xff);
      unsigned short v = 0;
      if ((t >> 24) && (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W)
        v = __ldg(img + ((int64_t)ih * p.W + iw) * p.C + ((t >> 16) & 0xff));
      if (e & 1) w[e >> 1] |= (uint32_t)v << 16;
      else w[e >> 1] = v;
    }
    *reinterpret_cast<uint4*>(p.out + (int64_t)m * p.Kp + j * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

}  // namespace im2col
}  // namespace mxb

using namespace mxb;

// x: bf16 [N,H,W,C] contiguous; out: bf16 [N*OH*OW, Kp], Kp = round_up(KH*KW*C, 8) (16-byte aligned rows).
MXB_API int MXBIm2colNHWC(const void* x, void* out, int N, int H, int W, int C, int KH, int KW, int stride, int pad,
                          int OH, int OW, int Kp, void* stream) {
  MXB_TRY
  const int K = KH * KW * C;
  MXB_CHECK(Kp % 8 == 0 && Kp >= K && Kp <= im2col::kMaxK, "im2col: Kp must be a multiple of 8, >= KH*KW*C and <= 512");
  MXB_CHECK(KH < 256 && KW < 256 && C < 256, "im2col: kernel extent / channels must be < 256");
  im2col::Params p;
  p.x = (const __nv_bfloat16*)x; p.out = (__nv_bfloat16*)out;
  p.H = H; p.W = W; p.C = C; p.KH = KH; p.KW = KW; p.stride = stride; p.pad = pad; p.OH = OH; p.OW = OW;
  p.K = K; p.Kp = Kp; p.chunks = Kp / 8;
  const int64_t M = (int64_t)N * OH * OW;
  p.total = M * p.chunks;
  MXB_CHECK(p.total < (1ll << 32), "im2col: too many output chunks for 32-bit indexing");
  if (p.total == 0) return 0;
  p.d_chunks = im2col::make_div((uint32_t)p.chunks);
  p.d_ohow = im2col::make_div((uint32_t)(OH * OW));
  p.d_ow = im2col::make_div((uint32_t)OW);
  const int grid = grid_for(p.total, im2col::kThreads, 2, 16);
  im2col::im2col_kernel<<<grid, im2col::kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}
