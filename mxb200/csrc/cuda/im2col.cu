// Patch extraction (im2col) for small-channel stem convolutions on channels-last bf16 images.
//
// Why: the ResNet stem (7x7 / stride 2 / 3 input channels; reference: Convolution op -> cuDNN,
// src/operator/nn/convolution.cu, falling back to im2col + GEMM, nn/im2col.cuh:79-180) has too few input channels
// for the tensor-core implicit-GEMM kernels of the library: measured on B200 at batch 256 the library takes 1.48 ms
// forward + 0.77 ms wgrad, 10-20x off the memory roofline of that layer.  Here the stem becomes
//     A[M = N*OH*OW, Kp] = patches (k = (kh*KWp + kw)*4 + c: channels padded to 4, kernel width to even)
//     Y = A . Wp^T      (our tcgen05 GEMM, BatchNorm statistics in its epilogue)
//     dW = dY^T . A     (same GEMM, TN mode, fp32 split-K reduce straight into the gradient arena)
// A is written once (16-byte stores, fully coalesced: consecutive threads own consecutive 16-byte chunks of a row)
// and kept for the backward pass; the input image is read through L1/L2 (each pixel is used by ~12 patches).
#include "common.cuh"

namespace mxb {
namespace im2col {

constexpr int kThreads = 256;
constexpr int kTileW = 64;      // output columns per block
constexpr int kCp = 4;          // channels are padded to 4 in the patch (8-byte pixels)
constexpr int kMaxStagePx = 2048;

struct Params {
  const __nv_bfloat16* x;   // [N, H, W, C], C <= 4
  __nv_bfloat16* out;       // [M, Kp], Kp = KH * KWp * 4
  int H, W, C, KH, KW, KWp, stride, pad, OH, OW, Kp, chunks;
};

// Patch layout: k = (kh * KWp + kw) * 4 + c with KWp = KW rounded up to even and the channel padded to 4, so a pixel is
// 8 bytes and every 16-byte chunk of a patch row is exactly two neighbouring input pixels.  One block = one (image,
// output row, tile of kTileW output columns):
//   phase 1 stages the KH input rows the tile touches in shared memory as 8-byte pixels (zero outside the image / for
//           the padded channel and column);
//   phase 2 writes the patch rows: one 16-byte chunk = two 8-byte shared loads + one 16-byte global store (~8
//           instructions per 8 output elements; the 2-byte gather version needed ~25 per element pair and was
//           instruction bound at 0.65-0.77 ms for the ResNet stem at batch 256).
__global__ void __launch_bounds__(kThreads) im2col_kernel(const Params p) {
  __shared__ uint2 stage[kMaxStagePx];
  const int tiles_w = (p.OW + kTileW - 1) / kTileW;
  const int tile = blockIdx.x % tiles_w;
  const int noh = blockIdx.x / tiles_w;
  const int oh = noh % p.OH, n = noh / p.OH;
  const int ow0 = tile * kTileW;
  const int tw = min(kTileW, p.OW - ow0);
  const int rowpix = (kTileW - 1) * p.stride + p.KWp;            // staged pixels per kernel row
  const int ih0 = oh * p.stride - p.pad, iw0 = ow0 * p.stride - p.pad;
  const unsigned short* img = reinterpret_cast<const unsigned short*>(p.x) + (int64_t)n * p.H * p.W * p.C;
  for (int i = threadIdx.x; i < p.KH * rowpix; i += kThreads) {
    const int kh = i / rowpix, px = i - kh * rowpix;
    const int ih = ih0 + kh, iw = iw0 + px;
    uint32_t c01 = 0, c23 = 0;
    if ((unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W) {
      const unsigned short* q = img + ((int64_t)ih * p.W + iw) * p.C;
      c01 = __ldg(q);
      if (p.C > 1) c01 |= (uint32_t)__ldg(q + 1) << 16;
      if (p.C > 2) c23 = __ldg(q + 2);
      if (p.C > 3) c23 |= (uint32_t)__ldg(q + 3) << 16;
    }
    stage[i] = make_uint2(c01, c23);
  }
  __syncthreads();
  const int rows_per_pass = kThreads / p.chunks;
  const int j = threadIdx.x % p.chunks, t0 = threadIdx.x / p.chunks;
  if (t0 >= rows_per_pass) return;
  const int half = p.KWp / 2;                      // 16-byte chunks per kernel row
  const int kh = j / half, jj = j - kh * half;
  const bool second_valid = (2 * jj + 1) < p.KW;   // the padded column (kw == KW) reads as zero
  const bool first_valid = (2 * jj) < p.KW;
  const int base = kh * rowpix + 2 * jj;
  const int64_t m0 = ((int64_t)n * p.OH + oh) * p.OW + ow0;
  for (int t = t0; t < tw; t += rows_per_pass) {
    const uint2 a = first_valid ? stage[base + t * p.stride] : make_uint2(0, 0);
    const uint2 b = second_valid ? stage[base + t * p.stride + 1] : make_uint2(0, 0);
    *reinterpret_cast<uint4*>(p.out + (m0 + t) * p.Kp + j * 8) = make_uint4(a.x, a.y, b.x, b.y);
  }
}

}  // namespace im2col
}  // namespace mxb

using namespace mxb;

// x: bf16 [N,H,W,C] contiguous, C <= 4; out: bf16 [N*OH*OW, KH*KWp*4] with KWp = KW rounded up to even.
MXB_API int MXBIm2colNHWC(const void* x, void* out, int N, int H, int W, int C, int KH, int KW, int stride, int pad,
                          int OH, int OW, int KWp, void* stream) {
  MXB_TRY
  MXB_CHECK(C >= 1 && C <= im2col::kCp, "im2col: at most 4 input channels");
  MXB_CHECK(KWp >= KW && KWp % 2 == 0 && KWp <= KW + 1, "im2col: KWp must be KW rounded up to even");
  im2col::Params p;
  p.x = (const __nv_bfloat16*)x; p.out = (__nv_bfloat16*)out;
  p.H = H; p.W = W; p.C = C; p.KH = KH; p.KW = KW; p.KWp = KWp; p.stride = stride; p.pad = pad; p.OH = OH; p.OW = OW;
  p.Kp = KH * KWp * im2col::kCp;
  p.chunks = p.Kp / 8;
  MXB_CHECK(p.chunks <= im2col::kThreads, "im2col: patch too long");
  MXB_CHECK(KH * ((im2col::kTileW - 1) * stride + KWp) <= im2col::kMaxStagePx, "im2col: kernel window too large for the staging buffer");
  const int64_t blocks = (int64_t)N * OH * ((OW + im2col::kTileW - 1) / im2col::kTileW);
  if (blocks == 0) return 0;
  MXB_CHECK(blocks < (1ll << 31), "im2col: grid too large");
  im2col::im2col_kernel<<<(unsigned)blocks, im2col::kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}
