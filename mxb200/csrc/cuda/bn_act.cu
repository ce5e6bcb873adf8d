// Fused BatchNorm (+ residual add) (+ ReLU) for channels-last activations, training forward + backward.
//
// Reference ops being fused: BatchNorm (src/operator/nn/batch_norm.cu:147-576, cuDNN path
// cudnn_batch_norm-inl.h), Activation(relu) (nn/activation.cu) and the residual elemwise_add -- three
// separate engine ops (and 4-6 full passes over the activation) in the reference.  Here:
//   forward : pass 1 reads x once (per-channel sum / sum-of-squares, fp32; the last block to finish
//             finalises mean / biased var / inv-std, updates the running statistics with
//             running = running*momentum + batch*(1-momentum) (batch_norm.cu:359-362) and emits the folded
//             scale/shift), pass 2 reads x (+residual) and writes y = relu(x*scale + shift + residual).
//             When the producer is our tcgen05 GEMM the sums come from its epilogue and pass 1 is skipped
//             (bn_finalize only).
//   backward: pass 1 reads dy, x (and y for the ReLU mask) once for sum(dy'), sum(dy'*xhat);
//             pass 2 writes dx = gamma*invstd*(dy' - mean(dy') - xhat*mean(dy'*xhat)) and, for residual
//             blocks, the masked gradient dy' for the skip branch in the same pass.
// Layout: x is [M, C] with C contiguous (NHWC); C*sizeof(T) must be a multiple of 16 and C/VEC a power of
// two <= 256 (all ResNet widths).  Statistics are always fp32 (AccReal in the reference).
#include <stdlib.h>

#include "common.cuh"

namespace mxb {
namespace bn {

constexpr int kThreads = 256;
constexpr int kSlabCV = 64;  // reduce kernels: a block covers at most 64 vector-columns (1 KB of a bf16 row); wider layers
                             // are split into column slabs so that parallelism (blocks) and RED traffic (row-blocks x 2C) decouple
constexpr int kUnroll = 4;   // independent 16-byte loads in flight per thread and per stream

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

struct Finalize {
  float* sum;          // [C] workspace (zeroed on exit)
  float* sumsq;        // [C]
  unsigned int* counter;  // arrival counter (zeroed on exit)
  float* partials;     // [grid][2C] per-block partial sums (no atomics; reduced by the last block)
  int use_atomics;     // wide layers: grid*2C partials would make the last block the bottleneck -> RED.ADD instead
  const float* gamma;  // may be null (=> 1)
  const float* beta;   // may be null (=> 0)
  float* running_mean; // may be null
  float* running_var;
  float* mean_out;     // [C] saved batch mean
  float* invstd_out;   // [C] saved 1/sqrt(var+eps)
  float* scale_out;    // [C] gamma*invstd
  float* shift_out;    // [C] beta - mean*gamma*invstd
  float eps, momentum;
  int64_t count;       // rows reduced (N*H*W)
};

__device__ __forceinline__ void finalize_channel(const Finalize& f, int c, float sum, float sumsq) {
  const float inv_n = 1.f / (float)f.count;
  const float mean = sum * inv_n;
  float var = sumsq * inv_n - mean * mean;
  var = fmaxf(var, 0.f);
  const float invstd = rsqrtf(var + f.eps);
  const float g = f.gamma ? f.gamma[c] : 1.f;
  const float b = f.beta ? f.beta[c] : 0.f;
  f.mean_out[c] = mean;
  f.invstd_out[c] = invstd;
  f.scale_out[c] = g * invstd;
  f.shift_out[c] = b - mean * g * invstd;
  if (f.running_mean) {
    f.running_mean[c] = f.running_mean[c] * f.momentum + mean * (1.f - f.momentum);
    f.running_var[c] = f.running_var[c] * f.momentum + var * (1.f - f.momentum);
  }
  f.sum[c] = 0.f;
  f.sumsq[c] = 0.f;
}

// Column-vector mapping: thread t owns vector column (t % CV) and rows (t / CV) + k * (kThreads / CV).
template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
stats_kernel(const T* __restrict__ x, int64_t M, int C, Finalize f) {
  constexpr int N = V<T>::N;
  const int CV = C / N;
  const int CVB = CV < kSlabCV ? CV : kSlabCV;   // vector-columns owned by this block (blockIdx.y = column slab)
  const int rpi = kThreads / CVB;                // rows per block iteration
  const int col = blockIdx.y * CVB + threadIdx.x % CVB;
  const int r0 = threadIdx.x / CVB;
  float s[N], s2[N];
#pragma unroll
  for (int j = 0; j < N; ++j) s[j] = s2[j] = 0.f;
  const int64_t rows_per_block = (M + gridDim.x - 1) / gridDim.x;
  const int64_t row_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t row_end = min(M, row_begin + rows_per_block);
  const T* base = x + (int64_t)col * N;
  int64_t r = row_begin + r0;
  // kUnroll independent 16B loads in flight per thread: HBM3e needs ~40 KB in flight per SM to stream at full rate
  for (; r + (kUnroll - 1) * rpi < row_end; r += kUnroll * rpi) {
    uint4 u[kUnroll];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) u[k] = ldg16_stream(base + (r + k * rpi) * C);
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
      float v[N];
      unpack<T>(u[k], v);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        s[j] += v[j];
        s2[j] = fmaf(v[j], v[j], s2[j]);
      }
    }
  }
  for (; r < row_end; r += rpi) {
    uint4 u0 = ldg16_stream(base + r * C);
    float v0[N];
    unpack<T>(u0, v0);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      s[j] += v0[j];
      s2[j] = fmaf(v0[j], v0[j], s2[j]);
    }
  }
  __shared__ float red[kThreads][2 * 8 + 1];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    red[threadIdx.x][j] = s[j];
    red[threadIdx.x][N + j] = s2[j];
  }
  __syncthreads();
  if (threadIdx.x < CVB) {
    float* mine = f.partials + (size_t)blockIdx.x * 2 * C;
#pragma unroll
    for (int j = 0; j < 2 * N; ++j) {
      float a = 0.f;
      for (int k = 0; k < rpi; ++k) a += red[threadIdx.x + k * CVB][j];
      const int idx = (j < N) ? (col * N + j) : (C + col * N + (j - N));
      if (f.use_atomics) atomicAdd(f.sum + idx, a);   // sum | sumsq are contiguous: [0,C) and [C,2C)
      else mine[idx] = a;
    }
  }
  // last block reduces the per-block partials (coalesced across channels, deterministic order) and finalises
  // publish this block's sums, then take a ticket: ONE fence by the ticket thread after the block barrier (fences are
  // cumulative over everything ordered before them by bar.sync); a fence in every thread cost 15 % of the kernel (ncu)
  __shared__ bool is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int prev = atomicAdd(f.counter, 1u);
    is_last = (prev == gridDim.x * gridDim.y - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    for (int c = threadIdx.x; c < C; c += kThreads) {
      float a = 0.f, b = 0.f;
      if (f.use_atomics) {
        a = __ldcg(f.sum + c);
        b = __ldcg(f.sumsq + c);
      } else {
        for (unsigned int k = 0; k < gridDim.x; ++k) {
          a += f.partials[(size_t)k * 2 * C + c];
          b += f.partials[(size_t)k * 2 * C + C + c];
        }
      }
      finalize_channel(f, c, a, b);
    }
    if (threadIdx.x == 0) *f.counter = 0u;
  }
}

__global__ void finalize_kernel(int C, Finalize f) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) finalize_channel(f, c, f.sum[c], f.sumsq[c]);
}

template <typename T, bool RELU, bool RES>
__global__ void __launch_bounds__(kThreads, 2)
apply_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y, uint8_t* __restrict__ mask,
             int64_t M, int C, const float* __restrict__ scale, const float* __restrict__ shift) {
  constexpr int N = V<T>::N;
  const int CV = C / N;
  const int64_t total = M * CV;
  const int64_t stride = (int64_t)gridDim.x * kThreads;  // multiple of CV
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int col = (int)(i % CV);
  float sc[N], sh[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    sc[j] = __ldg(scale + col * N + j);
    sh[j] = __ldg(shift + col * N + j);
  }
  auto body = [&](int64_t idx, const uint4& u, const uint4& ur) {
    float v[N];
    unpack<T>(u, v);
    if (RES) {
      float rv[N];
      unpack<T>(ur, rv);
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = fmaf(v[j], sc[j], sh[j]) + rv[j];
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
    }
    if (RELU) {
      uint32_t bits = 0;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        bits |= (v[j] > 0.f ? 1u : 0u) << j;
        v[j] = fmaxf(v[j], 0.f);
      }
      if (mask != nullptr) mask[idx] = (uint8_t)bits;   // 1 bit / element: what the backward needs instead of y
    }
    stg16(y + idx * N, pack<T>(v));
  };
  for (; i + (kUnroll - 1) * stride < total; i += kUnroll * stride) {
    uint4 u[kUnroll], ur[kUnroll];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
      u[k] = ldg16_stream(x + (i + k * stride) * N);
      if (RES) ur[k] = ldg16_stream(res + (i + k * stride) * N);
    }
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) body(i + k * stride, u[k], ur[k]);
  }
  for (; i < total; i += stride) {
    uint4 u = ldg16_stream(x + i * N), ur = make_uint4(0, 0, 0, 0);
    if (RES) ur = ldg16_stream(res + i * N);
    body(i, u, ur);
  }
}

struct BwdFinalize {
  float* sum_dy;       // [C] workspace, zeroed on exit
  float* sum_dyx;      // [C]
  unsigned int* counter;
  float* partials;     // [grid][2C]
  const float* gamma;  // may be null
  const float* invstd;
  float* dgamma;       // [C] out (may be null)
  float* dbeta;        // [C] out (may be null)
  float* c1;           // gamma*invstd
  float* c2;           // mean(dy')
  float* c3;           // mean(dy'*xhat)
  int64_t count;
  int accumulate;      // grad_req == 'add' for dgamma/dbeta
  int use_atomics;     // see Finalize::use_atomics
};

template <typename T, bool RELU>
__global__ void __launch_bounds__(kThreads, 2)
bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x, const uint8_t* __restrict__ mask, int64_t M,
                  int C, const float* __restrict__ mean, BwdFinalize f) {
  constexpr int N = V<T>::N;
  const int CV = C / N;
  const int CVB = CV < kSlabCV ? CV : kSlabCV;
  const int rpi = kThreads / CVB;
  const int col = blockIdx.y * CVB + threadIdx.x % CVB;
  const int r0 = threadIdx.x / CVB;
  float mu[N], is[N], s[N], s2[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    mu[j] = __ldg(mean + col * N + j);
    is[j] = __ldg(f.invstd + col * N + j);
    s[j] = s2[j] = 0.f;
  }
  const int64_t rows_per_block = (M + gridDim.x - 1) / gridDim.x;
  const int64_t row_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t row_end = min(M, row_begin + rows_per_block);
  auto body = [&](const uint4& ud, const uint4& ux, uint32_t bits) {
    float dv[N], xv[N];
    unpack<T>(ud, dv);
    unpack<T>(ux, xv);
    if (RELU) {
#pragma unroll
      for (int j = 0; j < N; ++j) dv[j] = ((bits >> j) & 1u) ? dv[j] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
      s[j] += dv[j];
      s2[j] = fmaf(dv[j], xv[j] - mu[j], s2[j]);   // invstd is applied once after the loop
    }
  };
  int64_t r = row_begin + r0;
  for (; r + (kUnroll - 1) * rpi < row_end; r += kUnroll * rpi) {
    uint4 ud[kUnroll], ux[kUnroll];
    uint32_t bits[kUnroll];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
      const int64_t off = (r + k * rpi) * C + (int64_t)col * N;
      ud[k] = ldg16_stream(dy + off);
      ux[k] = ldg16_stream(x + off);
      bits[k] = RELU ? (uint32_t)__ldg(mask + (r + k * rpi) * CV + col) : 0u;
    }
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) body(ud[k], ux[k], bits[k]);
  }
  for (; r < row_end; r += rpi) {
    const int64_t off = r * C + (int64_t)col * N;
    const uint4 ud = ldg16_stream(dy + off), ux = ldg16_stream(x + off);
    body(ud, ux, RELU ? (uint32_t)__ldg(mask + r * CV + col) : 0u);
  }
#pragma unroll
  for (int j = 0; j < N; ++j) s2[j] *= is[j];
  __shared__ float red[kThreads][2 * 8 + 1];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    red[threadIdx.x][j] = s[j];
    red[threadIdx.x][N + j] = s2[j];
  }
  __syncthreads();
  if (threadIdx.x < CVB) {
    float* mine = f.partials + (size_t)blockIdx.x * 2 * C;
#pragma unroll
    for (int j = 0; j < 2 * N; ++j) {
      float a = 0.f;
      for (int k = 0; k < rpi; ++k) a += red[threadIdx.x + k * CVB][j];
      const int idx = (j < N) ? (col * N + j) : (C + col * N + (j - N));
      if (f.use_atomics) atomicAdd(f.sum_dy + idx, a);   // sum_dy | sum_dyx are contiguous
      else mine[idx] = a;
    }
  }
  // publish this block's sums, then take a ticket: ONE fence by the ticket thread after the block barrier (fences are
  // cumulative over everything ordered before them by bar.sync); a fence in every thread cost 15 % of the kernel (ncu)
  __shared__ bool is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int prev = atomicAdd(f.counter, 1u);
    is_last = (prev == gridDim.x * gridDim.y - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    const float inv_n = 1.f / (float)f.count;
    for (int c = threadIdx.x; c < C; c += kThreads) {
      float sd = 0.f, sdx = 0.f;
      if (f.use_atomics) {
        sd = __ldcg(f.sum_dy + c);
        sdx = __ldcg(f.sum_dyx + c);
      } else {
        for (unsigned int k = 0; k < gridDim.x; ++k) {
          sd += f.partials[(size_t)k * 2 * C + c];
          sdx += f.partials[(size_t)k * 2 * C + C + c];
        }
      }
      const float g = f.gamma ? f.gamma[c] : 1.f;
      if (f.dgamma) f.dgamma[c] = (f.accumulate ? f.dgamma[c] : 0.f) + sdx;
      if (f.dbeta) f.dbeta[c] = (f.accumulate ? f.dbeta[c] : 0.f) + sd;
      f.c1[c] = g * f.invstd[c];
      f.c2[c] = sd * inv_n;
      f.c3[c] = sdx * inv_n;
      f.sum_dy[c] = 0.f;
      f.sum_dyx[c] = 0.f;
    }
    if (threadIdx.x == 0) *f.counter = 0u;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Backward reduce, bulk-copy pipelined (bf16).  The register-prefetch kernel above reaches ~2.5 TB/s (38 % of the measured
// copy bandwidth: the loads of a warp are in flight only while it is not computing, and every block pays a reduction
// tail after ~10 iterations).  Here a producer warp streams dy / x / mask through a 3-stage shared-memory ring with 1-D
// bulk copies (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier), so ~100 KB per SM stay in flight independent of
// the consumers, and each persistent block owns many chunks (one reduction tail per block, one wave of blocks).
//   chunk = 1024 16-byte vectors = R = 1024 / CV consecutive rows (16 KB of dy, 16 KB of x, 1 KB of mask)
//   consumer thread t owns vector column t % CV (CV is a power of two <= 256) -> 16 register accumulators
constexpr int kRedStages = 3;
constexpr int kRedVecs = 1024;                       // vectors per chunk and tensor
constexpr int kRedConsumers = 256;
constexpr int kRedThreads = kRedConsumers + 32;      // + producer warp
constexpr int kRedStageBytes = 2 * kRedVecs * 16 + kRedVecs;
constexpr int kRedSmem = kRedStages * kRedStageBytes + 128 + 1024;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(dst)),
               "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_init_(uint64_t* bar, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(n));
}
__device__ __forceinline__ void mbar_expect_(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity)
        : "memory");
    if (!ok && ++spins > (1u << 26)) __trap();
  }
}

template <bool RELU>
__global__ void __launch_bounds__(kRedThreads, 2)
bwd_reduce_bulk_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                       const uint8_t* __restrict__ mask, int64_t M, int C, const float* __restrict__ mean, BwdFinalize f) {
  extern __shared__ uint8_t red_smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(red_smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + kRedStages * kRedStageBytes);
  uint64_t* empty = full + kRedStages;
  const int CV = C / 8;
  const int R = kRedVecs / CV;                           // rows per chunk
  const int64_t nchunks = (M + R - 1) / R;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRedStages; ++i) {
      mbar_init_(&full[i], 1);
      mbar_init_(&empty[i], kRedConsumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float s[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = s2[j] = 0.f;
  const int col = threadIdx.x % CV;                      // consumers: fixed vector column (CV divides 256)
  if (warp == kRedConsumers / 32) {
    // ------------------------------------------------ producer: one thread streams this block's chunks
    if ((threadIdx.x & 31) == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const int64_t r0 = c * R;
        const int rows = (int)min((int64_t)R, M - r0);
        const uint32_t bytes = (uint32_t)rows * (uint32_t)C * 2u;
        const uint32_t mbytes = RELU ? (uint32_t)rows * (uint32_t)CV : 0u;
        mbar_wait_(&empty[stage], phase ^ 1);
        uint8_t* dst = sm + stage * kRedStageBytes;
        mbar_expect_(&full[stage], 2 * bytes + mbytes);
        bulk_g2s(dst, dy + r0 * C, bytes, &full[stage]);
        bulk_g2s(dst + kRedVecs * 16, x + r0 * C, bytes, &full[stage]);
        if (RELU) bulk_g2s(dst + 2 * kRedVecs * 16, mask + r0 * CV, mbytes, &full[stage]);
        if (++stage == kRedStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------ consumers
    float mu[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) mu[j] = __ldg(mean + col * 8 + j);
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
      const int64_t r0 = c * R;
      const int nvec = (int)min((int64_t)R, M - r0) * CV;
      mbar_wait_(&full[stage], phase);
      const uint4* sdy = reinterpret_cast<const uint4*>(sm + stage * kRedStageBytes);
      const uint4* sx = sdy + kRedVecs;
      const uint8_t* smk = reinterpret_cast<const uint8_t*>(sx + kRedVecs);
#pragma unroll
      for (int k = 0; k < kRedVecs / kRedConsumers; ++k) {
        const int v = threadIdx.x + k * kRedConsumers;
        if (v < nvec) {
          const uint4 ud = sdy[v], ux = sx[v];
          const uint32_t bits = RELU ? (uint32_t)smk[v] : 0xFFu;
          float dv[8], xv[8];
          unpack<__nv_bfloat16>(ud, dv);
          unpack<__nv_bfloat16>(ux, xv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float d = ((bits >> j) & 1u) ? dv[j] : 0.f;
            s[j] += d;
            s2[j] = fmaf(d, xv[j] - mu[j], s2[j]);
          }
        }
      }
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive_(&empty[stage]);
      if (++stage == kRedStages) {
        stage = 0;
        phase ^= 1;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s2[j] *= __ldg(f.invstd + col * 8 + j);
  }
  // ---------------------------------------------------- block reduction over the 256 / CV threads of a column, publish
  __syncthreads();                                        // every consumer is done with the ring: reuse it as scratch
  float (*red)[17] = reinterpret_cast<float (*)[17]>(sm);
  if (threadIdx.x < kRedConsumers) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[threadIdx.x][j] = s[j];
      red[threadIdx.x][8 + j] = s2[j];
    }
  }
  __syncthreads();
  if (threadIdx.x < CV) {
    const int reps = kRedConsumers / CV;
    float* mine = f.partials + (size_t)blockIdx.x * 2 * C;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float a = 0.f;
      for (int k = 0; k < reps; ++k) a += red[threadIdx.x + k * CV][j];
      const int idx = (j < 8) ? (col * 8 + j) : (C + col * 8 + (j - 8));
      if (f.use_atomics) atomicAdd(f.sum_dy + idx, a);
      else mine[idx] = a;
    }
  }
  __shared__ bool is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int prev = atomicAdd(f.counter, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    const float inv_n = 1.f / (float)f.count;
    for (int c = threadIdx.x; c < C; c += kRedThreads) {
      float sd = 0.f, sdx = 0.f;
      if (f.use_atomics) {
        sd = __ldcg(f.sum_dy + c);
        sdx = __ldcg(f.sum_dyx + c);
      } else {
        for (unsigned int k = 0; k < gridDim.x; ++k) {
          sd += f.partials[(size_t)k * 2 * C + c];
          sdx += f.partials[(size_t)k * 2 * C + C + c];
        }
      }
      const float g = f.gamma ? f.gamma[c] : 1.f;
      if (f.dgamma) f.dgamma[c] = (f.accumulate ? f.dgamma[c] : 0.f) + sdx;
      if (f.dbeta) f.dbeta[c] = (f.accumulate ? f.dbeta[c] : 0.f) + sd;
      f.c1[c] = g * f.invstd[c];
      f.c2[c] = sd * inv_n;
      f.c3[c] = sdx * inv_n;
      f.sum_dy[c] = 0.f;
      f.sum_dyx[c] = 0.f;
    }
    if (threadIdx.x == 0) *f.counter = 0u;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Elementwise BatchNorm passes on the same bulk-copy ring as bwd_reduce_bulk_kernel: inputs stream global -> shared memory
// through a producer warp (UBLKCP), the 256 consumer threads read their vectors from shared memory, compute and write
// 16-byte results straight to global memory (a warp writes 512 contiguous bytes).
//   MODE 0: forward apply   y = act(x * scale + shift (+ res)), 1-bit ReLU mask            in: x (, res)      out: y, mask
//   MODE 1: backward apply  dx = a1 * dy' + k1 * x + k0, dres = dy' (dy' = dy . mask)       in: dy, x, mask    out: dx (, dres)
struct ElemArgs {
  const __nv_bfloat16* in0;     // x  | dy
  const __nv_bfloat16* in1;     // res | x
  const uint8_t* mask_in;       // backward: ReLU mask
  __nv_bfloat16* out0;          // y  | dx
  __nv_bfloat16* out1;          // -  | dres
  uint8_t* mask_out;            // forward: ReLU mask
  const float* p0;              // scale | c1
  const float* p1;              // shift | c2
  const float* p2;              // -     | c3
  const float* mean;            // backward
  const float* invstd;          // backward
  int64_t M;
  int C;
};

template <int MODE, bool RELU, bool SECOND>     // SECOND: forward residual input / backward dres output
__global__ void __launch_bounds__(kRedThreads, 2)
elem_bulk_kernel(const ElemArgs a) {
  extern __shared__ uint8_t red_smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(red_smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + kRedStages * kRedStageBytes);
  uint64_t* empty = full + kRedStages;
  const int C = a.C, CV = C / 8;
  const int R = kRedVecs / CV;
  const int64_t nchunks = (a.M + R - 1) / R;
  constexpr bool kTwoIn = (MODE == 1) || SECOND;           // second 16 KB input stream
  constexpr bool kMaskIn = (MODE == 1) && RELU;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRedStages; ++i) {
      mbar_init_(&full[i], 1);
      mbar_init_(&empty[i], kRedConsumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if ((threadIdx.x >> 5) == kRedConsumers / 32) {
    if ((threadIdx.x & 31) == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const int64_t r0 = c * R;
        const int rows = (int)min((int64_t)R, a.M - r0);
        const uint32_t bytes = (uint32_t)rows * (uint32_t)C * 2u;
        const uint32_t mbytes = kMaskIn ? (uint32_t)rows * (uint32_t)CV : 0u;
        mbar_wait_(&empty[stage], phase ^ 1);
        uint8_t* dst = sm + stage * kRedStageBytes;
        mbar_expect_(&full[stage], (kTwoIn ? 2u : 1u) * bytes + mbytes);
        bulk_g2s(dst, a.in0 + r0 * C, bytes, &full[stage]);
        if (kTwoIn) bulk_g2s(dst + kRedVecs * 16, a.in1 + r0 * C, bytes, &full[stage]);
        if (kMaskIn) bulk_g2s(dst + 2 * kRedVecs * 16, a.mask_in + r0 * CV, mbytes, &full[stage]);
        if (++stage == kRedStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    return;
  }
  const int col = threadIdx.x % CV;
  float k1[8], k0[8], a1[8];      // forward: k1 = scale, k0 = shift;  backward: dx = a1 * dy' + k1 * x + k0
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (MODE == 0) {
      k1[j] = __ldg(a.p0 + col * 8 + j);
      k0[j] = __ldg(a.p1 + col * 8 + j);
      a1[j] = 0.f;
    } else {
      const float c1 = __ldg(a.p0 + col * 8 + j), c2 = __ldg(a.p1 + col * 8 + j), c3 = __ldg(a.p2 + col * 8 + j);
      const float mu = __ldg(a.mean + col * 8 + j), is = __ldg(a.invstd + col * 8 + j);
      a1[j] = c1;
      k1[j] = -c1 * is * c3;
      k0[j] = -k1[j] * mu - c1 * c2;
    }
  }
  int stage = 0;
  uint32_t phase = 0;
  for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int64_t r0 = c * R;
    const int nvec = (int)min((int64_t)R, a.M - r0) * CV;
    const int64_t v0 = r0 * CV;
    mbar_wait_(&full[stage], phase);
    const uint4* s0 = reinterpret_cast<const uint4*>(sm + stage * kRedStageBytes);
    const uint4* s1 = s0 + kRedVecs;
    const uint8_t* smk = reinterpret_cast<const uint8_t*>(s1 + kRedVecs);
    uint4 u0[4], u1[4];
    uint32_t bits[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int v = threadIdx.x + k * kRedConsumers;
      if (v < nvec) {
        u0[k] = s0[v];
        if (kTwoIn) u1[k] = s1[v];
        bits[k] = kMaskIn ? (uint32_t)smk[v] : 0xFFu;
      }
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive_(&empty[stage]);       // the ring slot is free as soon as it sits in registers
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int v = threadIdx.x + k * kRedConsumers;
      if (v >= nvec) continue;
      float x0[8], x1[8], o[8];
      unpack<__nv_bfloat16>(u0[k], x0);
      if (kTwoIn) unpack<__nv_bfloat16>(u1[k], x1);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(x0[j], k1[j], k0[j]) + (SECOND ? x1[j] : 0.f);
        if (RELU) {
          uint32_t m = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            m |= (o[j] > 0.f ? 1u : 0u) << j;
            o[j] = fmaxf(o[j], 0.f);
          }
          if (a.mask_out != nullptr) a.mask_out[v0 + v] = (uint8_t)m;
        }
        stg16(a.out0 + (v0 + v) * 8, pack<__nv_bfloat16>(o));
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = ((bits[k] >> j) & 1u) ? x0[j] : 0.f;
          x0[j] = d;
          o[j] = fmaf(a1[j], d, fmaf(k1[j], x1[j], k0[j]));
        }
        stg16(a.out0 + (v0 + v) * 8, pack<__nv_bfloat16>(o));
        if (SECOND) stg16(a.out1 + (v0 + v) * 8, pack<__nv_bfloat16>(x0));
      }
    }
    if (++stage == kRedStages) {
      stage = 0;
      phase ^= 1;
    }
  }
}

template <int MODE, bool RELU, bool SECOND>
static void launch_elem_bulk(const ElemArgs& a, cudaStream_t s) {
  const int CV = a.C / 8, R = kRedVecs / CV;
  const int64_t nchunks = (a.M + R - 1) / R;
  const int grid = (int)std::min<int64_t>(nchunks, (int64_t)kNumSMs * 2);
  auto kern = elem_bulk_kernel<MODE, RELU, SECOND>;
  CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kRedSmem));
  kern<<<grid, kRedThreads, kRedSmem, s>>>(a);
}

static bool elem_bulk_ok(int64_t M, int C, const void* p0, const void* p1, const void* p2, const void* p3, const void* mask) {
  static const bool off = [] { const char* e = getenv("MXB200_BN_BULK"); return e && e[0] == '0'; }();
  const int CV = C / 8;
  if (off || C % 8 != 0 || CV > 256 || (kRedVecs % CV) != 0) return false;
  const int R = kRedVecs / CV;
  auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % 16) == 0; };
  return M >= 4 * R && ((M * CV) % 16) == 0 && al(p0) && al(p1) && al(p2) && al(p3) && al(mask);
}

template <typename T, bool RELU, bool RES>
__global__ void __launch_bounds__(kThreads, 2)
bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, const uint8_t* __restrict__ mask, T* __restrict__ dx,
                 T* __restrict__ dres, int64_t M, int C, const float* __restrict__ mean,
                 const float* __restrict__ invstd, const float* __restrict__ c1, const float* __restrict__ c2,
                 const float* __restrict__ c3) {
  constexpr int N = V<T>::N;
  const int CV = C / N;
  const int64_t total = M * CV;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int col = (int)(i % CV);
  float mu[N], is[N], a1[N], a2[N], a3[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    mu[j] = __ldg(mean + col * N + j);
    is[j] = __ldg(invstd + col * N + j);
    a1[j] = __ldg(c1 + col * N + j);
    a2[j] = __ldg(c2 + col * N + j);
    a3[j] = __ldg(c3 + col * N + j);
  }
  float k1[N], k0[N];   // dx = a1*dv - a1*is*a3 * x + (a1*is*a3*mu - a1*a2)  -> two FMAs per element
#pragma unroll
  for (int j = 0; j < N; ++j) {
    k1[j] = -a1[j] * is[j] * a3[j];
    k0[j] = -k1[j] * mu[j] - a1[j] * a2[j];
  }
  auto body = [&](int64_t idx, const uint4& ud, const uint4& ux, uint32_t bits) {
    float dv[N], xv[N], o[N];
    unpack<T>(ud, dv);
    unpack<T>(ux, xv);
    if (RELU) {
#pragma unroll
      for (int j = 0; j < N; ++j) dv[j] = ((bits >> j) & 1u) ? dv[j] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) o[j] = fmaf(a1[j], dv[j], fmaf(k1[j], xv[j], k0[j]));
    stg16(dx + idx * N, pack<T>(o));
    if (RES) stg16(dres + idx * N, pack<T>(dv));
  };
  for (; i + (kUnroll - 1) * stride < total; i += kUnroll * stride) {
    uint4 ud[kUnroll], ux[kUnroll];
    uint32_t bits[kUnroll];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
      ud[k] = ldg16_stream(dy + (i + k * stride) * N);
      ux[k] = ldg16_stream(x + (i + k * stride) * N);
      bits[k] = RELU ? (uint32_t)__ldg(mask + i + k * stride) : 0u;
    }
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) body(i + k * stride, ud[k], ux[k], bits[k]);
  }
  for (; i < total; i += stride) {
    const uint4 ud = ldg16_stream(dy + i * N), ux = ldg16_stream(x + i * N);
    body(i, ud, ux, RELU ? (uint32_t)__ldg(mask + i) : 0u);
  }
}

template <typename T>
static dim3 grid_rows(int64_t M, int C) {
  const int CV = C / V<T>::N;
  const int CVB = CV < kSlabCV ? CV : kSlabCV;
  const int slabs = CV / CVB;
  const int rpi = kThreads / CVB;
  // one full wave of 4 resident blocks per SM in total; each block should own at least ~2 unrolled iterations
  int64_t gx = (M + (int64_t)rpi * kUnroll * 2 - 1) / ((int64_t)rpi * kUnroll * 2);
  const int64_t cap = std::max<int64_t>(1, (int64_t)kNumSMs * 4 / slabs);
  gx = std::max<int64_t>(1, std::min<int64_t>(gx, cap));
  return dim3((unsigned)gx, (unsigned)slabs, 1);
}

static int grid_elems(int64_t total_vec, int CV) {
  int64_t blocks = (total_vec + kThreads * 4 - 1) / (kThreads * 4);
  blocks = std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)kNumSMs * 8));
  (void)CV;
  return (int)blocks;
}

template <typename T>
static void check_shape(int C) {
  constexpr int N = V<T>::N;
  if (C % N != 0) throw std::runtime_error("bn: C must be a multiple of the 16B vector width");
  const int CV = C / N;
  if (CV > kThreads || (CV & (CV - 1)) != 0) throw std::runtime_error("bn: C/vec must be a power of two <= 256");
}

}  // namespace bn
}  // namespace mxb

using namespace mxb;
using namespace mxb::bn;

// ws: fp32 workspace with a layout that does NOT depend on C (calls with different widths share it):
//   ws[0]                      arrival counter (uint32)
//   ws[8 .. 8+2C)              column sums | sums of squares (also where a producer GEMM deposits its statistics)
//   ws[8+2*kMaxC .. )          per-block partials, kMaxGrid * 2C floats
// ws[0 .. 8+2C) is left zeroed on exit.
static constexpr int kMaxGrid = kNumSMs * 4;
static constexpr int kMaxC = 4096;
static constexpr int kSumOff = 8;
static constexpr int kPartOff = 8 + 2 * kMaxC;

// a single block has to sum grid * 2C partials: fine up to a few hundred KB, a bottleneck beyond (measured: 0.6 ms
// for C=2048) -> wide layers combine their block sums with fp32 RED.ADD instead (order-nondeterministic in the
// last bits; MXB200_BN_DETERMINISTIC=1 forces the partials path)
static int use_atomics_for(int C) {
  static const bool det = [] { const char* e = getenv("MXB200_BN_DETERMINISTIC"); return e && e[0] == '1'; }();
  return (!det && C >= 256) ? 1 : 0;
}

MXB_API int MXBBnWorkspaceFloats(int C) { return kPartOff + kMaxGrid * 2 * (C > kMaxC ? kMaxC : C); }

MXB_API int MXBBnStatsFinalize(const void* x, int dt, int64_t M, int C, float* ws, const float* gamma,
                               const float* beta, float* running_mean, float* running_var, float* mean_out,
                               float* invstd_out, float* scale_out, float* shift_out, float eps, float momentum,
                               int stats_precomputed, void* stream) {
  MXB_TRY
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  MXB_CHECK(C <= kMaxC, "bn: at most 4096 channels");
  Finalize f{ws + kSumOff, ws + kSumOff + C, reinterpret_cast<unsigned int*>(ws), ws + kPartOff, use_atomics_for(C), gamma, beta,
             running_mean, running_var, mean_out, invstd_out, scale_out, shift_out, eps, momentum, M};
  if (stats_precomputed) {
    finalize_kernel<<<(C + 255) / 256, 256, 0, s>>>(C, f);
  } else if (dt == kBF16) {
    check_shape<__nv_bfloat16>(C);
    stats_kernel<__nv_bfloat16><<<grid_rows<__nv_bfloat16>(M, C), kThreads, 0, s>>>((const __nv_bfloat16*)x, M, C, f);
  } else if (dt == kF32) {
    check_shape<float>(C);
    stats_kernel<float><<<grid_rows<float>(M, C), kThreads, 0, s>>>((const float*)x, M, C, f);
  } else {
    throw std::runtime_error("bn: dtype must be bf16 or fp32");
  }
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}

// mask: [M * C/VEC] bytes (bit j of byte i = (y > 0) of element j of vector i); may be null.
MXB_API int MXBBnApply(const void* x, const void* res, void* y, void* mask, int dt, int64_t M, int C,
                       const float* scale, const float* shift, int relu, void* stream) {
  MXB_TRY
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  auto go = [&](auto t) {
    using T = decltype(t);
    check_shape<T>(C);
    const int CV = C / V<T>::N;
    const int g = grid_elems(M * CV, CV);
    const T* xx = (const T*)x; const T* rr = (const T*)res; T* yy = (T*)y; uint8_t* mm = (uint8_t*)mask;
    if constexpr (sizeof(T) == 2) {
      if (elem_bulk_ok(M, C, x, res, y, nullptr, nullptr)) {
        ElemArgs ea{(const __nv_bfloat16*)x, (const __nv_bfloat16*)res, nullptr, (__nv_bfloat16*)y, nullptr, mm, scale, shift,
                    nullptr, nullptr, nullptr, M, C};
        if (relu && res) launch_elem_bulk<0, true, true>(ea, s);
        else if (relu) launch_elem_bulk<0, true, false>(ea, s);
        else if (res) launch_elem_bulk<0, false, true>(ea, s);
        else launch_elem_bulk<0, false, false>(ea, s);
        return;
      }
    }
    if (relu && res) apply_kernel<T, true, true><<<g, kThreads, 0, s>>>(xx, rr, yy, mm, M, C, scale, shift);
    else if (relu) apply_kernel<T, true, false><<<g, kThreads, 0, s>>>(xx, rr, yy, mm, M, C, scale, shift);
    else if (res) apply_kernel<T, false, true><<<g, kThreads, 0, s>>>(xx, rr, yy, mm, M, C, scale, shift);
    else apply_kernel<T, false, false><<<g, kThreads, 0, s>>>(xx, rr, yy, mm, M, C, scale, shift);
  };
  if (dt == kBF16) go(__nv_bfloat16{});
  else if (dt == kF32) go(float{});
  else throw std::runtime_error("bn: dtype must be bf16 or fp32");
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}

// ws as above.  coef: 3*C floats scratch (c1,c2,c3).  relu => mask must be the byte map written by MXBBnApply.
// dgamma / dbeta: fp32, written (accumulate=0) or accumulated into (accumulate=1; e.g. straight into a gradient arena).
MXB_API int MXBBnBackward(const void* dy, const void* x, const void* mask, void* dx, void* dres, int dt, int64_t M,
                          int C, const float* mean, const float* invstd, const float* gamma, float* dgamma,
                          float* dbeta, float* ws, float* coef, int relu, int accumulate, void* stream) {
  MXB_TRY
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  MXB_CHECK(C <= kMaxC, "bn: at most 4096 channels");
  BwdFinalize f{ws + kSumOff, ws + kSumOff + C, reinterpret_cast<unsigned int*>(ws), ws + kPartOff, gamma, invstd, dgamma,
                dbeta, coef, coef + C, coef + 2 * C, M, accumulate, use_atomics_for(C)};
  auto go = [&](auto t) {
    using T = decltype(t);
    check_shape<T>(C);
    const int CV = C / V<T>::N;
    const T* dd = (const T*)dy; const T* xx = (const T*)x; const uint8_t* mm = (const uint8_t*)mask;
    static const bool bulk_off = [] { const char* e = getenv("MXB200_BN_BULK"); return e && e[0] == '0'; }();
    bool bulk = false;
    if constexpr (sizeof(T) == 2) {
      // alignment of every chunk start: rows * C * 2 and rows * CV bytes must be multiples of 16
      const int R = kRedVecs / CV;
      bulk = !bulk_off && CV <= 256 && (kRedVecs % CV) == 0 && M >= 4 * R && ((M * CV) % 16) == 0 &&
             (reinterpret_cast<uintptr_t>(dy) % 16) == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0 &&
             (!relu || reinterpret_cast<uintptr_t>(mask) % 16 == 0);
      if (bulk) {
        const int64_t nchunks = (M + R - 1) / R;
        int grid = (int)std::min<int64_t>(nchunks, (int64_t)kNumSMs * 2);
        if (grid > kMaxGrid) grid = kMaxGrid;
        // Small layers (< 300 MB of input): the streaming part takes 8-30 us and a fixed ~40 us appears on top once the
        // grid goes beyond one CTA per SM (profiles/step_roofline_bn_r2_call32.txt: 25088 x 512 ran at 1.0 TB/s,
        // 100352 x 256 at 2.0 TB/s).  Measured A/B (profiles/bn_reduce_grid_ab_r2_call33.txt): 103 MB 53 -> 36 us at
        // 148 CTAs, 51 MB 50 -> 29 us at 111 CTAs, 205 MB 56 -> 53 us at 185 CTAs; larger inputs need the full grid (a
        // CTA's 96 KB ring streams ~24 GB/s).  Pick the grid that minimises
        //   t(grid) = max(bytes / HBM, bytes / (grid * 24 GB/s)) + 30 us * (grid - 148) / 148   [second term only above 148]
        // and combine the block sums with RED.ADD.  MXB200_BN_REDUCE_GRID=0: always 2 CTAs per SM.
        static const bool grid_model = [] { const char* e = getenv("MXB200_BN_REDUCE_GRID"); return !(e && e[0] == '0'); }();
        static const bool det = [] { const char* e = getenv("MXB200_BN_DETERMINISTIC"); return e && e[0] == '1'; }();
        const double bytes = (double)M * C * 4.0;
        if (grid_model && !det && bytes < 300e6) {
          f.use_atomics = 1;
          double best = 1e30;
          int best_grid = grid;
          for (int g = kNumSMs / 2; g <= kNumSMs * 2; g += kNumSMs / 4) {
            if (g > nchunks || g > kMaxGrid) break;
            const double stream = std::max(bytes / 6.4e12, bytes / (g * 24e9));
            const double t = stream + (g > kNumSMs ? 30e-6 * (g - kNumSMs) / kNumSMs : 0.0);
            if (t < best) {
              best = t;
              best_grid = g;
            }
          }
          grid = best_grid;
        }
        const __nv_bfloat16* d16 = reinterpret_cast<const __nv_bfloat16*>(dy);
        const __nv_bfloat16* x16 = reinterpret_cast<const __nv_bfloat16*>(x);
        if (relu) {
          CUDA_CHECK(cudaFuncSetAttribute(bwd_reduce_bulk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRedSmem));
          bwd_reduce_bulk_kernel<true><<<grid, kRedThreads, kRedSmem, s>>>(d16, x16, mm, M, C, mean, f);
        } else {
          CUDA_CHECK(cudaFuncSetAttribute(bwd_reduce_bulk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRedSmem));
          bwd_reduce_bulk_kernel<false><<<grid, kRedThreads, kRedSmem, s>>>(d16, x16, mm, M, C, mean, f);
        }
      }
    }
    if (bulk) {
    } else if (relu) bwd_reduce_kernel<T, true><<<grid_rows<T>(M, C), kThreads, 0, s>>>(dd, xx, mm, M, C, mean, f);
    else bwd_reduce_kernel<T, false><<<grid_rows<T>(M, C), kThreads, 0, s>>>(dd, xx, mm, M, C, mean, f);
    const int g = grid_elems(M * CV, CV);
    T* ox = (T*)dx; T* orr = (T*)dres;
    if constexpr (sizeof(T) == 2) {
      if (elem_bulk_ok(M, C, dy, x, dx, dres, relu ? mask : nullptr)) {
        ElemArgs ea{(const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, mm, (__nv_bfloat16*)dx, (__nv_bfloat16*)dres, nullptr,
                    f.c1, f.c2, f.c3, mean, invstd, M, C};
        if (relu && dres) launch_elem_bulk<1, true, true>(ea, s);
        else if (relu) launch_elem_bulk<1, true, false>(ea, s);
        else if (dres) launch_elem_bulk<1, false, true>(ea, s);
        else launch_elem_bulk<1, false, false>(ea, s);
        return;
      }
    }
    if (relu && dres) bwd_apply_kernel<T, true, true><<<g, kThreads, 0, s>>>(dd, xx, mm, ox, orr, M, C, mean, invstd, f.c1, f.c2, f.c3);
    else if (relu) bwd_apply_kernel<T, true, false><<<g, kThreads, 0, s>>>(dd, xx, mm, ox, orr, M, C, mean, invstd, f.c1, f.c2, f.c3);
    else if (dres) bwd_apply_kernel<T, false, true><<<g, kThreads, 0, s>>>(dd, xx, mm, ox, orr, M, C, mean, invstd, f.c1, f.c2, f.c3);
    else bwd_apply_kernel<T, false, false><<<g, kThreads, 0, s>>>(dd, xx, mm, ox, orr, M, C, mean, invstd, f.c1, f.c2, f.c3);
  };
  if (dt == kBF16) go(__nv_bfloat16{});
  else if (dt == kF32) go(float{});
  else throw std::runtime_error("bn: dtype must be bf16 or fp32");
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}
