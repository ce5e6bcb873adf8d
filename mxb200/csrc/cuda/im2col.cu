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

// One block = one (image, output row, tile of kTileW output columns).  Phase 1 stages the KH input rows that the tile
// touches in shared memory (coalesced 2-byte loads, zero fill outside the image), phase 2 writes the patch rows with
// 16-byte stores: thread -> fixed 16-byte chunk j of the patch (its 8 source offsets live in registers) and output
// columns t0, t0 + rows_per_pass, ...  The input is read from global memory once per tile instead of ~KH*KW/stride^2
// times through L1.
constexpr int kTileW = 64;
constexpr int kMaxStage = 8192;   // staged input elements per block (KH * ((kTileW-1)*stride + KW) * C)

__global__ void __launch_bounds__(kThreads) im2col_kernel(const Params p) {
  __shared__ unsigned short stage[kMaxStage];
  const int tiles_w = (p.OW + kTileW - 1) / kTileW;
  const int tile = blockIdx.x % tiles_w;
  const int noh = blockIdx.x / tiles_w;
  const int oh = noh % p.OH, n = noh / p.OH;
  const int ow0 = tile * kTileW;
  const int tw = min(kTileW, p.OW - ow0);
  const int span = ((kTileW - 1) * p.stride + p.KW) * p.C;       // staged elements per kernel row
  const int ih0 = oh * p.stride - p.pad, iw0 = ow0 * p.stride - p.pad;
  const unsigned short* img = reinterpret_cast<const unsigned short*>(p.x) + (int64_t)n * p.H * p.W * p.C;
  for (int i = threadIdx.x; i < p.KH * span; i += kThreads) {
    const int kh = i / span, e = i - kh * span;
    const int ih = ih0 + kh, iw = iw0 + e / p.C;
    unsigned short v = 0;
    if ((unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W) v = __ldg(img + ((int64_t)ih * p.W) * p.C + (int64_t)iw0 * p.C + e);
    stage[i] = v;
  }
  __syncthreads();
  const int rows_per_pass = kThreads / p.chunks;
  const int j = threadIdx.x % p.chunks, t0 = threadIdx.x / p.chunks;
  if (t0 >= rows_per_pass) return;
  int off[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = j * 8 + e;
    if (k < p.K) {
      const int kh = k / (p.KW * p.C);
      off[e] = kh * span + (k - kh * p.KW * p.C);
    } else {
      off[e] = -1;
    }
  }
  const int64_t m0 = ((int64_t)n * p.OH + oh) * p.OW + ow0;
  for (int t = t0; t < tw; t += rows_per_pass) {
    const int base = t * p.stride * p.C;
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const uint32_t v = off[e] >= 0 ? stage[off[e] + base] : 0u;
      if (e & 1) w[e >> 1] |= v << 16;
      else w[e >> 1] = v;
    }
    *reinterpret_cast<uint4*>(p.out + (m0 + t) * p.Kp + j * 8) = make_uint4(w[0], w[1], w[2], w[3]);
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
  MXB_CHECK(KH * (((im2col::kTileW - 1) * stride + KW) * C) <= im2col::kMaxStage, "im2col: kernel window too large for the staging buffer");
  MXB_CHECK(p.chunks <= im2col::kThreads, "im2col: patch too long");
  const int64_t blocks = (int64_t)N * OH * ((OW + im2col::kTileW - 1) / im2col::kTileW);
  MXB_CHECK(blocks < (1ll << 31), "im2col: grid too large");
  im2col::im2col_kernel<<<(unsigned)blocks, im2col::kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}
