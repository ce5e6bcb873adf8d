// LayerNorm over the last axis (training forward + backward) and fused elementwise activations
// (bias + GELU/… forward and backward) for transformer blocks.
//
// Reference ops: LayerNorm (src/operator/nn/layer_norm-inl.h: mean/std via broadcast-reduce kernels, then three
// more elementwise passes; backward is ~8 generic reduce/elementwise launches), LeakyReLU/Activation
// (src/operator/leaky_relu-inl.h, nn/activation-inl.h: one mshadow expression per call).
// Here: one kernel per direction.
//   forward : one warp per row; the row lives in registers (16-byte vector loads, VPT vectors per lane), mean and
//             variance by warp shuffles (two-pass in registers: no E[x^2]-E[x]^2 cancellation), saves mean / rstd.
//   backward: same mapping; dx in one pass; dgamma/dbeta are accumulated per thread over the rows its warp owns,
//             folded across the block's warps in shared memory and written as per-block partials that a second
//             tiny kernel sums (deterministic, no atomics), optionally accumulating into an fp32 gradient arena.
#include "common.cuh"

namespace mxb {
namespace ln {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

template <typename T>
__device__ __forceinline__ void unpack8(const uint4& u, float* o);
template <>
__device__ __forceinline__ void unpack8<__nv_bfloat16>(const uint4& u, float* o) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[2 * i] = __low2float(p[i]);
    o[2 * i + 1] = __high2float(p[i]);
  }
}
template <typename T>
__device__ __forceinline__ uint4 pack8(const float* o);
template <>
__device__ __forceinline__ uint4 pack8<__nv_bfloat16>(const float* o) {
  uint4 u;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
  return u;
}

// fp32 rows are handled as two 16-byte vectors per "8-element vector" so both dtypes share the mapping:
// lane l owns elements [ (j*32 + l)*8, +8 ) for j < VPT  (C = VPT*256).
template <typename T>
__device__ __forceinline__ void load8(const T* p, float* o);
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float* o) {
  unpack8<__nv_bfloat16>(ldg16(p), o);
}
template <>
__device__ __forceinline__ void load8<float>(const float* p, float* o) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float* o);
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float* o) {
  stg16(p, pack8<__nv_bfloat16>(o));
}
template <>
__device__ __forceinline__ void store8<float>(float* p, const float* o) {
  *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(o[4], o[5], o[6], o[7]);
}

template <typename T, int VPT>
__global__ void __launch_bounds__(kThreads)
fwd_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ sum_out, const float* __restrict__ gamma,
           const float* __restrict__ beta, T* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
           int64_t rows, float eps) {
  constexpr int C = VPT * 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float g[VPT][8], b[VPT][8];
#pragma unroll
  for (int j = 0; j < VPT; ++j)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = (j * 32 + lane) * 8 + k;
      g[j][k] = gamma ? __ldg(gamma + c) : 1.f;
      b[j][k] = beta ? __ldg(beta + c) : 0.f;
    }
  for (int64_t r = (int64_t)blockIdx.x * kWarps + warp; r < rows; r += (int64_t)gridDim.x * kWarps) {
    float v[VPT][8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      load8<T>(x + r * C + (j * 32 + lane) * 8, v[j]);
      if (res != nullptr) {   // y = LN(x + res): the sum is rounded to T and stored (it is what backward needs)
        float rv[8];
        load8<T>(res + r * C + (j * 32 + lane) * 8, rv);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[j][k] = to_f(from_f<T>(v[j][k] + rv[k]));
        store8<T>(sum_out + r * C + (j * 32 + lane) * 8, v[j]);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[j][k];
    }
    const float mean = warp_sum(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d = v[j][k] - mean;
        q = fmaf(d, d, q);
      }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
    if (lane == 0) {
      mean_out[r] = mean;
      rstd_out[r] = rstd;
    }
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = fmaf((v[j][k] - mean) * rstd, g[j][k], b[j][k]);
      store8<T>(y + r * C + (j * 32 + lane) * 8, o);
    }
  }
}

template <typename T, int VPT>
__global__ void __launch_bounds__(kThreads)
bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ gamma,
           const float* __restrict__ mean, const float* __restrict__ rstd, T* __restrict__ dx,
           float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows) {
  constexpr int C = VPT * 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float g[VPT][8], dg[VPT][8], db[VPT][8];
#pragma unroll
  for (int j = 0; j < VPT; ++j)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      g[j][k] = gamma ? __ldg(gamma + (j * 32 + lane) * 8 + k) : 1.f;
      dg[j][k] = db[j][k] = 0.f;
    }
  for (int64_t r = (int64_t)blockIdx.x * kWarps + warp; r < rows; r += (int64_t)gridDim.x * kWarps) {
    const float mu = mean[r], rs = rstd[r];
    float xh[VPT][8], gy[VPT][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      float xv[8], dv[8];
      load8<T>(x + r * C + (j * 32 + lane) * 8, xv);
      load8<T>(dy + r * C + (j * 32 + lane) * 8, dv);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        xh[j][k] = (xv[k] - mu) * rs;
        gy[j][k] = dv[k] * g[j][k];
        s1 += gy[j][k];
        s2 = fmaf(gy[j][k], xh[j][k], s2);
        dg[j][k] = fmaf(dv[k], xh[j][k], dg[j][k]);
        db[j][k] += dv[k];
      }
    }
    const float m1 = warp_sum(s1) * (1.f / C), m2 = warp_sum(s2) * (1.f / C);
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = rs * (gy[j][k] - m1 - xh[j][k] * m2);
      store8<T>(dx + r * C + (j * 32 + lane) * 8, o);
    }
  }
  // fold the 8 warps of the block (same lane -> same columns), then one fp32 RED per column and block into the
  // (accumulating) outputs -- a second "sum the per-block partials" kernel cost as much as this one (19 us, latency bound)
  __shared__ float red[kWarps][32 * 8 + 1];
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    float* dst = which ? dbeta : dgamma;
    if (dst == nullptr) continue;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 8; ++k) red[warp][lane * 8 + k] = which ? db[j][k] : dg[j][k];
      __syncthreads();
      for (int i = threadIdx.x; i < 256; i += kThreads) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) a += red[w][i];
        atomicAdd(dst + j * 256 + i, a);
      }
    }
  }
}

__global__ void reduce_partials_kernel(const float* __restrict__ partials, int grid, int C, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= 2 * C) return;
  float a = 0.f;
  for (int k = 0; k < grid; ++k) a += partials[(size_t)k * 2 * C + c];
  float* dst = c < C ? (dgamma ? dgamma + c : nullptr) : (dbeta ? dbeta + (c - C) : nullptr);
  if (dst) *dst = (accumulate ? *dst : 0.f) + a;
}

constexpr int kMaxGrid = kNumSMs * 2;

// ------------------------------------------------------------------ activations
enum Act : int { kRelu = 1, kGeluErf = 2, kGeluTanh = 3, kSigmoid = 4, kTanh = 5, kSoftRelu = 6, kLeaky = 7, kElu = 8,
                 kSelu = 9, kSilu = 10 };

__device__ __forceinline__ float act_fwd(int a, float x, float slope) {
  switch (a) {
    case kRelu: return fmaxf(x, 0.f);
    case kGeluErf: return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
    case kGeluTanh: {
      const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
      return 0.5f * x * (1.f + tanhf(u));
    }
    case kSigmoid: return 1.f / (1.f + __expf(-x));
    case kTanh: return tanhf(x);
    case kSoftRelu: return x > 20.f ? x : log1pf(__expf(x));
    case kLeaky: return x > 0.f ? x : slope * x;
    case kElu: return x > 0.f ? x : slope * (__expf(x) - 1.f);
    case kSelu: return 1.0507009873554805f * (x > 0.f ? x : 1.6732632423543772f * (__expf(x) - 1.f));
    case kSilu: return x / (1.f + __expf(-x));
    default: return x;
  }
}

// derivative w.r.t. the pre-activation x
__device__ __forceinline__ float act_bwd(int a, float x, float slope) {
  switch (a) {
    case kRelu: return x > 0.f ? 1.f : 0.f;
    case kGeluErf: {
      const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
      return cdf + x * 0.3989422804014327f * __expf(-0.5f * x * x);
    }
    case kGeluTanh: {
      const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
      const float t = tanhf(u);
      return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * 0.7978845608028654f * (1.f + 3.f * 0.044715f * x * x);
    }
    case kSigmoid: {
      const float s = 1.f / (1.f + __expf(-x));
      return s * (1.f - s);
    }
    case kTanh: {
      const float t = tanhf(x);
      return 1.f - t * t;
    }
    case kSoftRelu: return 1.f / (1.f + __expf(-x));
    case kLeaky: return x > 0.f ? 1.f : slope;
    case kElu: return x > 0.f ? 1.f : slope * __expf(x);
    case kSelu: return 1.0507009873554805f * (x > 0.f ? 1.f : 1.6732632423543772f * __expf(x));
    case kSilu: {
      const float s = 1.f / (1.f + __expf(-x));
      return s * (1.f + x * (1.f - s));
    }
    default: return 1.f;
  }
}

// y = act(x + bias[col]);  x: [rows, C] contiguous, C % 8 == 0 (bias may be null).  8 elements per thread-iteration.
template <typename T>
__global__ void __launch_bounds__(256)
bias_act_fwd_kernel(const T* __restrict__ x, const float* __restrict__ bias, T* __restrict__ y, int64_t total8, int C8,
                    int act, float slope) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total8; i += (int64_t)gridDim.x * 256) {
    float v[8];
    load8<T>(x + i * 8, v);
    const int c = (int)(i % C8) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = act_fwd(act, v[k] + (bias ? __ldg(bias + c + k) : 0.f), slope);
    store8<T>(y + i * 8, v);
  }
}

// dx = dy * act'(x + bias[col])
template <typename T>
__global__ void __launch_bounds__(256)
bias_act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ bias,
                    T* __restrict__ dx, int64_t total8, int C8, int act, float slope) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total8; i += (int64_t)gridDim.x * 256) {
    float v[8], d[8];
    load8<T>(x + i * 8, v);
    load8<T>(dy + i * 8, d);
    const int c = (int)(i % C8) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] *= act_bwd(act, v[k] + (bias ? __ldg(bias + c + k) : 0.f), slope);
    store8<T>(dx + i * 8, d);
  }
}

// out[c] += sum over rows of x[r, c]  (bias gradients of FullyConnected: db = colsum(dy); the reference runs a generic
// reduce, fully_connected-inl.h:162-171).  Block = 32 vector-columns x 8 row lanes, 4 independent loads in flight per
// thread, shared-memory fold over the row lanes, one fp32 RED per column and block.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, float* __restrict__ out, int64_t M, int C, int rows_per_block) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int vcol = blockIdx.y * 32 + tx;
  const bool active = vcol * 8 < C;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r_end = min(M, r_begin + rows_per_block);
  if (active) {
    int64_t r = r_begin + ty;
    for (; r + 24 < r_end; r += 32) {
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) load8<T>(x + (r + 8 * u) * C + vcol * 8, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += v[u][k];
    }
    for (; r < r_end; r += 8) {
      float v[8];
      load8<T>(x + r * C + vcol * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v[k];
    }
  }
  __shared__ float red[8][32 * 8 + 1];
#pragma unroll
  for (int k = 0; k < 8; ++k) red[ty][tx * 8 + k] = acc[k];
  __syncthreads();
  const int c = threadIdx.x;                       // 256 threads <-> 256 columns of this block
  const int col = blockIdx.y * 256 + c;
  if (col < C) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += red[w][c];
    atomicAdd(out + col, a);
  }
}

}  // namespace ln
}  // namespace mxb

using namespace mxb;
using namespace mxb::ln;

MXB_API int MXBLayerNormWorkspaceFloats(int C) { return kMaxGrid * 2 * C; }

template <typename T, int VPT>
static void ln_fwd(const void* x, const void* res, void* sum_out, const float* gamma, const float* beta, void* y,
                   float* mean, float* rstd, int64_t rows, float eps, cudaStream_t s) {
  int grid = (int)std::min<int64_t>((rows + kWarps - 1) / kWarps, (int64_t)kNumSMs * 8);
  fwd_kernel<T, VPT><<<grid, kThreads, 0, s>>>((const T*)x, (const T*)res, (T*)sum_out, gamma, beta, (T*)y, mean, rstd, rows, eps);
}

template <typename T, int VPT>
static void ln_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, void* dx,
                   float* dgamma, float* dbeta, float* ws, int64_t rows, int accumulate, cudaStream_t s) {
  constexpr int C = VPT * 256;
  (void)ws;
  int grid = (int)std::min<int64_t>((rows + kWarps * 4 - 1) / (kWarps * 4), (int64_t)kMaxGrid);
  if (grid < 1) grid = 1;
  if (!accumulate) {
    if (dgamma) CUDA_CHECK(cudaMemsetAsync(dgamma, 0, C * sizeof(float), s));
    if (dbeta) CUDA_CHECK(cudaMemsetAsync(dbeta, 0, C * sizeof(float), s));
  }
  bwd_kernel<T, VPT><<<grid, kThreads, 0, s>>>((const T*)dy, (const T*)x, gamma, mean, rstd, (T*)dx, dgamma, dbeta, rows);
}

#define MXB_LN_DISPATCH(FN, ...)                                                            \
  do {                                                                                      \
    const int vpt = C / 256;                                                                \
    if (dt == kBF16) {                                                                      \
      switch (vpt) {                                                                        \
        case 1: FN<__nv_bfloat16, 1>(__VA_ARGS__); break;                                   \
        case 2: FN<__nv_bfloat16, 2>(__VA_ARGS__); break;                                   \
        case 3: FN<__nv_bfloat16, 3>(__VA_ARGS__); break;                                   \
        case 4: FN<__nv_bfloat16, 4>(__VA_ARGS__); break;                                   \
        case 6: FN<__nv_bfloat16, 6>(__VA_ARGS__); break;                                   \
        case 8: FN<__nv_bfloat16, 8>(__VA_ARGS__); break;                                   \
        default: throw std::runtime_error("layernorm: unsupported width (C/256 must be 1,2,3,4,6,8)"); \
      }                                                                                     \
    } else if (dt == kF32) {                                                                \
      switch (vpt) {                                                                        \
        case 1: FN<float, 1>(__VA_ARGS__); break;                                           \
        case 2: FN<float, 2>(__VA_ARGS__); break;                                           \
        case 3: FN<float, 3>(__VA_ARGS__); break;                                           \
        case 4: FN<float, 4>(__VA_ARGS__); break;                                           \
        case 6: FN<float, 6>(__VA_ARGS__); break;                                           \
        case 8: FN<float, 8>(__VA_ARGS__); break;                                           \
        default: throw std::runtime_error("layernorm: unsupported width (C/256 must be 1,2,3,4,6,8)"); \
      }                                                                                     \
    } else {                                                                                \
      throw std::runtime_error("layernorm: dtype must be bf16 or fp32");                    \
    }                                                                                       \
  } while (0)

// x,y: [rows, C] contiguous, C % 256 == 0 and C/256 in {1,2,3,4,6,8}; gamma/beta fp32 [C] (may be null).
// res (optional, same shape): y = LN(x + res) and sum_out receives x + res (the tensor backward differentiates).
MXB_API int MXBLayerNormForward(const void* x, const void* res, void* sum_out, int dt, const float* gamma,
                                const float* beta, void* y, float* mean, float* rstd, int64_t rows, int C, float eps,
                                void* stream) {
  MXB_TRY
  MXB_CHECK(C % 256 == 0, "layernorm: C must be a multiple of 256");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  MXB_CHECK(res == nullptr || sum_out != nullptr, "layernorm: residual needs a sum output");
  MXB_LN_DISPATCH(ln_fwd, x, res, sum_out, gamma, beta, y, mean, rstd, rows, eps, s);
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}

// ws: MXBLayerNormWorkspaceFloats(C) floats.  dgamma/dbeta fp32 [C] (may be null); accumulate adds into them.
MXB_API int MXBLayerNormBackward(const void* dy, const void* x, int dt, const float* gamma, const float* mean,
                                 const float* rstd, void* dx, float* dgamma, float* dbeta, float* ws, int64_t rows,
                                 int C, int accumulate, void* stream) {
  MXB_TRY
  MXB_CHECK(C % 256 == 0, "layernorm: C must be a multiple of 256");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  MXB_LN_DISPATCH(ln_bwd, dy, x, gamma, mean, rstd, dx, dgamma, dbeta, ws, rows, accumulate, s);
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}

// y = act(x + bias); x,y [rows, C], C % 8 == 0.  act codes: see ln::Act.
MXB_API int MXBBiasActForward(const void* x, int dt, const float* bias, void* y, int64_t rows, int C, int act,
                              float slope, void* stream) {
  MXB_TRY
  MXB_CHECK(C % 8 == 0, "bias_act: C must be a multiple of 8");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t total8 = rows * (C / 8);
  if (total8 == 0) return 0;
  const int grid = grid_for(total8, 256, 2, 16);
  if (dt == kBF16) bias_act_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, bias, (__nv_bfloat16*)y, total8, C / 8, act, slope);
  else if (dt == kF32) bias_act_fwd_kernel<float><<<grid, 256, 0, s>>>((const float*)x, bias, (float*)y, total8, C / 8, act, slope);
  else throw std::runtime_error("bias_act: dtype must be bf16 or fp32");
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}

MXB_API int MXBBiasActBackward(const void* dy, const void* x, int dt, const float* bias, void* dx, int64_t rows, int C,
                               int act, float slope, void* stream) {
  MXB_TRY
  MXB_CHECK(C % 8 == 0, "bias_act: C must be a multiple of 8");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t total8 = rows * (C / 8);
  if (total8 == 0) return 0;
  const int grid = grid_for(total8, 256, 2, 16);
  if (dt == kBF16) bias_act_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, bias, (__nv_bfloat16*)dx, total8, C / 8, act, slope);
  else if (dt == kF32) bias_act_bwd_kernel<float><<<grid, 256, 0, s>>>((const float*)dy, (const float*)x, bias, (float*)dx, total8, C / 8, act, slope);
  else throw std::runtime_error("bias_act: dtype must be bf16 or fp32");
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}

// out (fp32 [C]) += column sums of x [rows, C]; C % 8 == 0.  Accumulates (zero ``out`` first for a plain sum).
MXB_API int MXBColSumAccumulate(const void* x, int dt, float* out, int64_t rows, int C, void* stream) {
  MXB_TRY
  MXB_CHECK(C % 8 == 0, "colsum: C must be a multiple of 8");
  if (rows == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int rpb = 128;
  while (rpb < 4096 && (rows + rpb - 1) / rpb * ((C + 255) / 256) > kNumSMs * 8) rpb *= 2;
  dim3 grid((unsigned)((rows + rpb - 1) / rpb), (unsigned)((C + 255) / 256), 1);
  if (dt == kBF16) colsum_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, out, rows, C, rpb);
  else if (dt == kF32) colsum_kernel<float><<<grid, 256, 0, s>>>((const float*)x, out, rows, C, rpb);
  else throw std::runtime_error("colsum: dtype must be bf16 or fp32");
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}
