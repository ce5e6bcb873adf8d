// Softmax / softmax-cross-entropy kernels.
//
// Reference: SoftmaxOutput (src/operator/softmax_output-inl.h:100-240; mshadow SoftmaxKernel /
// SoftmaxGradKernel, 3rdparty/mshadow/mshadow/cuda/tensor_gpu-inl.cuh:240-470) and the Gluon loss path
// log_softmax + pick + mean (python/mxnet/gluon/loss.py:243-320 -> nn/softmax-inl.h:166-268), i.e. 3-5
// kernels and two extra passes over the logits.  Here one block per row computes, in a single launch,
// the row max / log-sum-exp, the per-row loss AND the logits gradient
//     d = (softmax(x) - onehot(label) [with label smoothing]) * scale        (0 for ignored rows)
// with fp32 math on bf16/fp16/fp32 logits.
#include "common.cuh"

namespace mxb {
namespace sm {

constexpr int kThreads = 256;

__device__ __forceinline__ float block_reduce_max(float v, float* red) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = (threadIdx.x & 31) < kThreads / 32 ? red[threadIdx.x & 31] : -INFINITY;
  r = warp_max(r);  // every warp reduces the same partials -> identical result in all threads
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = (threadIdx.x & 31) < kThreads / 32 ? red[threadIdx.x & 31] : 0.f;
  r = warp_sum(r);
  __syncthreads();
  return r;
}

// mode 0: y = softmax(x); mode 1: y = log_softmax(x)
template <typename T>
__global__ void __launch_bounds__(kThreads)
softmax_rows_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t rows, int cols, int mode, float temperature) {
  __shared__ float red[kThreads / 32];
  const float it = 1.f / temperature;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const T* xr = x + r * cols;
    float m = -INFINITY;
    for (int c = threadIdx.x; c < cols; c += kThreads) m = fmaxf(m, to_f(xr[c]) * it);
    m = block_reduce_max(m, red);
    float s = 0.f;
    for (int c = threadIdx.x; c < cols; c += kThreads) s += __expf(to_f(xr[c]) * it - m);
    s = block_reduce_sum(s, red);
    const float inv = 1.f / s, ls = __logf(s);
    T* yr = y + r * cols;
    for (int c = threadIdx.x; c < cols; c += kThreads) {
      const float z = to_f(xr[c]) * it - m;
      yr[c] = from_f<T>(mode == 0 ? __expf(z) * inv : z - ls);
    }
  }
}

// Fused CE: loss[r] = lse(x_r) - (1-a)*x_r[l] - a/(K-1) * sum_{j!=l} x_r[j]   (a = smooth_alpha)
//           dx     = (softmax - target) * scale_r
// scale_r = grad_scale * (1/valid_count if norm_valid) where the valid count is supplied by the caller.
template <typename T, typename LT>
__global__ void __launch_bounds__(kThreads)
softmax_ce_kernel(const T* __restrict__ x, const LT* __restrict__ label, float* __restrict__ loss,
                  T* __restrict__ dx, int64_t rows, int cols, float grad_scale, int use_ignore, float ignore_label,
                  float smooth_alpha, const float* __restrict__ sample_weight) {
  __shared__ float red[kThreads / 32];
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const T* xr = x + r * cols;
    const float lf = (float)label[r];
    const int l = (int)lf;
    const bool ignored = use_ignore && lf == ignore_label;
    float m = -INFINITY, sx = 0.f;
    for (int c = threadIdx.x; c < cols; c += kThreads) {
      const float v = to_f(xr[c]);
      m = fmaxf(m, v);
      sx += v;
    }
    m = block_reduce_max(m, red);
    float s = 0.f;
    for (int c = threadIdx.x; c < cols; c += kThreads) s += __expf(to_f(xr[c]) - m);
    s = block_reduce_sum(s, red);
    const float lse = m + __logf(s);
    const float w = (sample_weight ? sample_weight[r] : 1.f);
    if (loss != nullptr) {
      if (smooth_alpha > 0.f) sx = block_reduce_sum(sx, red);
      if (threadIdx.x == 0) {
        float v = 0.f;
        if (!ignored && l >= 0 && l < cols) {
          const float xl = to_f(xr[l]);
          const float off = smooth_alpha > 0.f ? smooth_alpha / (float)(cols - 1) : 0.f;
          v = lse - (1.f - smooth_alpha) * xl - off * (sx - xl);
        }
        loss[r] = v * w;
      }
    }
    if (dx != nullptr) {
      T* dr = dx + r * cols;
      const float inv = 1.f / s;
      const float on = 1.f - smooth_alpha;
      const float off = smooth_alpha > 0.f ? smooth_alpha / (float)(cols - 1) : 0.f;
      const float sc = ignored ? 0.f : grad_scale * w;
      for (int c = threadIdx.x; c < cols; c += kThreads) {
        const float p = __expf(to_f(xr[c]) - m) * inv;
        dr[c] = from_f<T>((p - (c == l ? on : off)) * sc);
      }
    }
  }
}

// SoftmaxOutput backward from saved probabilities: g = (p - target) * scale (0 for ignored rows).
template <typename T, typename LT>
__global__ void __launch_bounds__(kThreads)
softmax_output_grad_kernel(const T* __restrict__ p, const LT* __restrict__ label, T* __restrict__ g, int64_t rows,
                           int cols, float scale, int use_ignore, float ignore_label, float smooth_alpha) {
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const float lf = (float)label[r];
    const int l = (int)lf;
    const bool ignored = use_ignore && lf == ignore_label;
    const float on = 1.f - smooth_alpha;
    const float off = smooth_alpha > 0.f ? smooth_alpha / (float)(cols - 1) : 0.f;
    for (int c = threadIdx.x; c < cols; c += kThreads) {
      const float v = ignored ? 0.f : (to_f(p[r * cols + c]) - (c == l ? on : off)) * scale;
      g[r * cols + c] = from_f<T>(v);
    }
  }
}

template <typename F>
static void dispatch(int dt, F&& f) {
  switch (dt) {
    case kF32: f(float{}); break;
    case kBF16: f(__nv_bfloat16{}); break;
    case kF16: f(__half{}); break;
    default: throw std::runtime_error("softmax: unsupported dtype");
  }
}

}  // namespace sm
}  // namespace mxb

using namespace mxb;
using namespace mxb::sm;

MXB_API int MXBSoftmaxRows(const void* x, void* y, int dt, int64_t rows, int cols, int mode, float temperature,
                           void* stream) {
  MXB_TRY
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (int)std::min<int64_t>(rows, (int64_t)kNumSMs * 8);
  dispatch(dt, [&](auto t) {
    using T = decltype(t);
    softmax_rows_kernel<T><<<grid, kThreads, 0, s>>>((const T*)x, (T*)y, rows, cols, mode, temperature);
  });
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}

// label_dt: 0 = float32 labels, 6 = int64 labels
MXB_API int MXBSoftmaxCE(const void* x, int dt, const void* label, int label_dt, float* loss, void* dx, int64_t rows,
                         int cols, float grad_scale, int use_ignore, float ignore_label, float smooth_alpha,
                         const float* sample_weight, void* stream) {
  MXB_TRY
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (int)std::min<int64_t>(rows, (int64_t)kNumSMs * 8);
  dispatch(dt, [&](auto t) {
    using T = decltype(t);
    if (label_dt == 6)
      softmax_ce_kernel<T, int64_t><<<grid, kThreads, 0, s>>>((const T*)x, (const int64_t*)label, loss, (T*)dx, rows,
                                                            cols, grad_scale, use_ignore, ignore_label, smooth_alpha,
                                                            sample_weight);
    else
      softmax_ce_kernel<T, float><<<grid, kThreads, 0, s>>>((const T*)x, (const float*)label, loss, (T*)dx, rows, cols,
                                                          grad_scale, use_ignore, ignore_label, smooth_alpha,
                                                          sample_weight);
  });
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}

MXB_API int MXBSoftmaxOutputGrad(const void* p, int dt, const void* label, int label_dt, void* g, int64_t rows,
                                 int cols, float scale, int use_ignore, float ignore_label, float smooth_alpha,
                                 void* stream) {
  MXB_TRY
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (int)std::min<int64_t>(rows, (int64_t)kNumSMs * 8);
  dispatch(dt, [&](auto t) {
    using T = decltype(t);
    if (label_dt == 6)
      softmax_output_grad_kernel<T, int64_t><<<grid, kThreads, 0, s>>>((const T*)p, (const int64_t*)label, (T*)g, rows,
                                                                     cols, scale, use_ignore, ignore_label,
                                                                     smooth_alpha);
    else
      softmax_output_grad_kernel<T, float><<<grid, kThreads, 0, s>>>((const T*)p, (const float*)label, (T*)g, rows,
                                                                   cols, scale, use_ignore, ignore_label, smooth_alpha);
  });
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}
