// kvstore('device') push+pull as ONE kernel: cross-GPU gradient reduction over NVLink peer memory fused
// with gradient rescale/clip, the optimizer update and the weight fan-out.
//
// What it replaces in the reference: CommDevice::Reduce = (N-1) cudaMemcpyPeerAsync +
// ElementwiseSum<gpu> + a blocking stream sync inside the engine op (src/kvstore/comm.h:503-550,
// src/ndarray/ndarray.cc:1318-1331), then the Python updater callback launching sgd_mom_update /
// mp_sgd_mom_update per tensor (optimizer_op-inl.h:305-448), then CommDevice::Broadcast = (N+1) copies
// (comm.h:598-616) -- per parameter tensor, ~161 times per ResNet-50 step; and KVStoreNCCL's
// ncclReduce + ncclBcast pair (kvstore_nccl.h:270-308,386-413).
//
// Here every GPU runs the same persistent kernel over a symmetric, peer-mapped arena that holds ALL
// parameters (chunk table gives per-tensor lr/wd multipliers):
//   * two-shot (bandwidth-bound sizes): rank r owns slice r of the arena.  It loads that slice of every
//     peer's gradient arena straight over NVLink (16B vector loads from peer pointers), sums in fp32,
//     applies the partitioned optimizer update (fp32 master weight + momentum exist only for the owned
//     slice: ZeRO-1-like state partitioning), and stores the updated low-precision weight slice into
//     EVERY rank's weight arena (peer stores).  Per-GPU NVLink traffic = 2*(N-1)/N*S.
//   * one-shot (latency-bound sizes): every rank loads the whole message from all peers and applies the
//     update redundantly on its own replica: one barrier pair, no second phase.
//   * flags in peer memory (st.release.sys / ld.acquire.sys, monotonically increasing sequence numbers,
//     never reset) replace cudaStreamSynchronize + engine ops.  Spins are bounded: on timeout the kernel
//     records an error in a host-visible word and exits instead of hanging the GPU.
// No NCCL call and no separate optimizer kernel runs on this path.
#include "common.cuh"
#include "optim_math.cuh"

namespace mxb {
namespace ar {

constexpr int kMaxRanks = 8;
constexpr int kMaxBlocks = 256;   // flag pad rows
constexpr int kThreads = 512;
constexpr int kChunk = 1024;
constexpr uint32_t kSpinLimit = 200u * 1000u * 1000u;

enum Mode : int {
  kAllreduce = 0,   // out_p[i] = scale * sum_q grad_q[i]            (push without updater + pull)
  kSgd = 1,         // SGD / momentum / NAG update, weights fanned out
  kAdam = 2,
  kLambPhase1 = 3,  // reduce + moments + direction r (local slice) + per-tensor partial norms
  kLambPhase2 = 4,  // w -= lr*trust*r on the slice, weights fanned out (no gradient traffic)
  kBroadcast = 5,   // weights_p[i] = src[i] for the owned slice (init / re-shard)
  kLamb = 6,        // phases 1 + 2 in one launch with an in-kernel cross-rank reduction of the per-tensor norms
};

struct Args {
  void* grads[kMaxRanks];     // gradient arenas as mapped in this process
  void* weights[kMaxRanks];   // weight (or allreduce output) arenas
  uint32_t* flags[kMaxRanks]; // signal pads: [kMaxBlocks][kMaxRanks] uint32
  float* w32;                 // local fp32 master arena (may be null: weight arena is the master)
  float* mom;                 // local momentum / Adam m
  float* var;                 // Adam / LAMB v
  float* rbuf;                // LAMB direction
  float* w_norm2;             // LAMB per-tensor partial norms (local)
  float* r_norm2;
  const int* chunk_tid;
  const float* lr_mult;
  const float* wd_mult;
  int* err;                   // host-visible error word
  int64_t n;                  // arena elements (multiple of kChunk)
  int rank, nranks;
  int mode, one_shot;
  int g_dt, w_dt;
  uint32_t seq;
  int zero_grad;              // zero the local gradient slice after it was consumed by everyone
  float lr, wd, momentum, rescale, clip;
  int nesterov, has_mom;
  float beta1, beta2, eps, bc1, bc2;
  float lamb_lower, lamb_upper;
  // CUDA-graph friendly hyper-parameters: when non-null the per-step scalars are read from this device buffer instead
  // of the by-value fields above (a captured launch freezes its by-value arguments).  Layout (floats):
  //   [0] lr  [1] wd  [2] momentum  [3] rescale  [4] clip  [5] seq (uint32 bits)  [6] bc1  [7] bc2
  // ``seq`` then acts as an offset added to hyper[5] (several launches of one step share one buffer).
  const float* hyper;
  // NVLS: multicast mappings of the gradient / weight arenas (null: P2P loads / stores over the unicast peer mappings)
  void* grads_mc;
  void* weights_mc;
  // fused LAMB (mode kLamb): per-rank partial norm buffers [2][nt_pad] in peer-mapped memory, local totals, grid barrier
  float* norms[kMaxRanks];
  float* norm_tot;
  uint32_t* gsync;            // local {count, generation} words of the intra-rank grid barrier
  int nt_pad;
  int pad0;                   // != 0: kU-fold unrolled gather in the bandwidth-bound modes
  int64_t state_off;
  // Device-resident barrier sequence number {seq, finished-CTA count}: when non-null the kernel reads seq from here and the
  // last CTA to finish increments it, so replaying a captured launch needs no host-side bookkeeping at all (every rank
  // advances its own copy identically: one increment per launch on the group).
  uint32_t* seq_dev;
  // Low-latency one-shot (small messages): every rank PUSHES its fp32 gradients as {value, sequence-flag} 8-byte packets into
  // every peer's receive buffer [2 parities][nranks][n] and then polls its own buffer -- no barrier at all, the flag travels
  // with the data (the idea of NCCL's LL protocol).  Null: flag-barrier protocol.
  uint2* ll_recv[kMaxRanks];
  uint32_t* ll_ctr;           // per-arena launch counter (device): packet flag and buffer parity; advanced with seq_dev
};
constexpr int kHyperFloats = 8;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Block-level barrier with the same block index on every peer.
__device__ __forceinline__ bool peer_barrier(const Args& a, uint32_t val) {
  __shared__ int ok;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  if (threadIdx.x < a.nranks) {
    const int p = threadIdx.x;
    st_release_sys(a.flags[p] + blockIdx.x * kMaxRanks + a.rank, val);
    const uint32_t* mine = a.flags[a.rank] + blockIdx.x * kMaxRanks + p;
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(mine) - val) < 0) {
      if (++spins > kSpinLimit) {
        ok = 0;
        if (a.err) atomicExch(a.err, 1 + p);
        break;
      }
    }
  }
  __syncthreads();
  return ok != 0;
}

template <typename T>
__device__ __forceinline__ void ld8(const T* p, float (&o)[8]);
template <>
__device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&o)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[2 * i] = __low2float(h[i]);
    o[2 * i + 1] = __high2float(h[i]);
  }
}
template <>
__device__ __forceinline__ void ld8<float>(const float* p, float (&o)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
template <typename T>
__device__ __forceinline__ void st8(T* p, const float (&o)[8]);
template <>
__device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, const float (&o)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
template <>
__device__ __forceinline__ void st8<float>(float* p, const float (&o)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(o[4], o[5], o[6], o[7]);
}

// ---- NVLS: in-switch reduction / broadcast through a multicast mapping of the symmetric arenas (csrc/cuda/vmm.cu).
// One multimem.ld_reduce returns the sum over ALL ranks' copies of 16 bytes (the NVSwitch reduces; SASS: LDGMC.ADD),
// one multimem.st writes 16 bytes into every rank's copy.  Per GPU the two-shot exchange then moves S(N-1)/N out + S/N in
// for the gradients instead of S(N-1)/N in only, i.e. both link directions carry payload.
template <typename T>
__device__ __forceinline__ void mc_ld_reduce8(const T* p, float (&o)[8]);
template <>
__device__ __forceinline__ void mc_ld_reduce8<float>(const float* p, float (&o)[8]) {
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]) : "l"(p) : "memory");
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(o[4]), "=f"(o[5]), "=f"(o[6]), "=f"(o[7]) : "l"(p + 4) : "memory");
}
template <>
__device__ __forceinline__ void mc_ld_reduce8<__nv_bfloat16>(const __nv_bfloat16* p, float (&o)[8]) {
  uint4 u;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p) : "memory");
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[2 * i] = __low2float(h[i]);
    o[2 * i + 1] = __high2float(h[i]);
  }
}
template <typename T>
__device__ __forceinline__ void mc_st8(T* p, const float (&o)[8]);
template <>
__device__ __forceinline__ void mc_st8<float>(float* p, const float (&o)[8]) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]) : "memory");
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p + 4), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7]) : "memory");
}
template <>
__device__ __forceinline__ void mc_st8<__nv_bfloat16>(__nv_bfloat16* p, const float (&o)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
  asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}

// All CTAs of this rank (co-resident by construction: <= 1 per SM).  count returns to zero, gen is monotonic.
__device__ __forceinline__ void grid_barrier(uint32_t* gsync) {
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile uint32_t* gen = gsync + 1;
    const uint32_t g = *gen;
    __threadfence();
    if (atomicAdd(gsync, 1u) == gridDim.x - 1) {
      gsync[0] = 0;
      __threadfence();
      atomicAdd(gsync + 1, 1u);
    } else {
      uint32_t spins = 0;
      while (*gen == g) {
        if (++spins > kSpinLimit) break;
      }
    }
    __threadfence();
  }
  __syncthreads();
}

template <typename GT, typename WT, int NR>
__global__ void __launch_bounds__(kThreads)
fused_kernel(const Args a) {
  const int nr = NR > 0 ? NR : a.nranks;
  // per-step scalars: by value, or from the device-resident hyper buffer (graph replay)
  float h_lr = a.lr, h_wd = a.wd, h_mom = a.momentum, h_rescale = a.rescale, h_clip = a.clip, h_bc1 = a.bc1, h_bc2 = a.bc2;
  uint32_t h_seq = a.seq;
  if (a.hyper != nullptr) {
    h_lr = __ldg(a.hyper + 0); h_wd = __ldg(a.hyper + 1); h_mom = __ldg(a.hyper + 2); h_rescale = __ldg(a.hyper + 3);
    h_clip = __ldg(a.hyper + 4); h_seq += __float_as_uint(__ldg(a.hyper + 5)); h_bc1 = __ldg(a.hyper + 6);
    h_bc2 = __ldg(a.hyper + 7);
  }
  if (a.seq_dev != nullptr) h_seq = a.seq + *reinterpret_cast<const volatile uint32_t*>(a.seq_dev);
  const bool ll = (sizeof(GT) == 4) && a.ll_recv[0] != nullptr && a.ll_ctr != nullptr && a.seq_dev != nullptr && nr > 1 && a.one_shot && a.mode != kLambPhase2 &&
                  a.mode != kBroadcast && a.mode != kLamb && a.mode != kLambPhase1;
  const uint32_t ll_cnt = ll ? *reinterpret_cast<const volatile uint32_t*>(a.ll_ctr) : 0u;
  const uint32_t ll_flag = ll_cnt | 0x80000000u;           // never 0 (buffers start zeroed)
  const int64_t ll_base = (int64_t)(ll_cnt & 1u) * nr * a.n;   // double buffering: a peer is never 2 launches of this arena behind
  if (ll) {
    if constexpr (sizeof(GT) == 4) {
      // push my gradients to every rank (my own buffer included: the reduction below treats all sources alike)
      const float* gl = reinterpret_cast<const float*>(a.grads[a.rank]);
      const int64_t nv2 = a.n / 2;
      for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < nv2; v += (int64_t)gridDim.x * kThreads) {
        const float2 g2 = *reinterpret_cast<const float2*>(gl + v * 2);
        const uint4 pkt = make_uint4(__float_as_uint(g2.x), ll_flag, __float_as_uint(g2.y), ll_flag);
#pragma unroll
        for (int k = 0; k < (NR > 0 ? NR : kMaxRanks); ++k) {
          if (k < nr) {
            const int p = (a.rank + k) % nr;
            uint4* dst = reinterpret_cast<uint4*>(a.ll_recv[p] + ll_base + (int64_t)a.rank * a.n) + v;
            asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(pkt.x), "r"(pkt.y), "r"(pkt.z), "r"(pkt.w) : "memory");
          }
        }
      }
    }
  } else if (nr > 1) {
    if (!peer_barrier(a, 4u * h_seq)) return;
  }
  // slice owned by this rank (two-shot) or the whole arena (one-shot)
  const int64_t n_chunks = a.n / kChunk;
  int64_t c_begin = 0, c_end = n_chunks;
  if (!a.one_shot && nr > 1) {
    const int64_t per = (n_chunks + nr - 1) / nr;
    c_begin = min(n_chunks, per * a.rank);
    c_end = min(n_chunks, c_begin + per);
  }
  const bool fan_out = !a.one_shot && nr > 1;
  const bool mc_grads = a.grads_mc != nullptr && nr > 1;
  const bool mc_weights = a.weights_mc != nullptr && fan_out;
  constexpr int kVecPerChunk = kChunk / 8;  // 128 vectors of 8 elements
  const int64_t v_begin = c_begin * kVecPerChunk, v_end = c_end * kVecPerChunk;
  const int64_t so = a.state_off;           // optimizer state arrays start at this arena element (slice-only state)
  __shared__ float red[2][kThreads / 32];

  auto fan = [&](int64_t e, const float (&w)[8]) {
    if (mc_weights) {
      mc_st8<WT>(reinterpret_cast<WT*>(a.weights_mc) + e, w);
    } else if (fan_out) {
#pragma unroll
      for (int k = 0; k < (NR > 0 ? NR : kMaxRanks); ++k)
        if (k < nr) st8<WT>(reinterpret_cast<WT*>(a.weights[(a.rank + k) % nr]) + e, w);
    } else {
      st8<WT>(reinterpret_cast<WT*>(a.weights[a.rank]) + e, w);
    }
  };

  // Per-vector metadata from the chunk table (tensor id -> lr / wd multipliers; padding chunks are skipped).
  struct Meta { int64_t e; int t; float lr, wd; bool skip; };
  auto meta_of = [&](int64_t v) {
    Meta m;
    const bool active = v < v_end;
    m.e = v * 8;
    m.t = 0; m.lr = h_lr; m.wd = h_wd; m.skip = !active;
    if (active && a.chunk_tid != nullptr) {
      m.t = __ldg(a.chunk_tid + m.e / kChunk);
      if (m.t < 0) m.skip = true;
      else {
        m.lr *= a.lr_mult ? __ldg(a.lr_mult + m.t) : 1.f;
        m.wd *= a.wd_mult ? __ldg(a.wd_mult + m.t) : 1.f;
      }
    }
    return m;
  };

  // Everything after the gradient of one 8-element vector is known: optimizer math, master / state update, weight fan-out.
  auto process = [&](const int mode, const Meta& mm, float (&g)[8], float* wn2, float* rn2) {
    const int64_t e = mm.e;
    const int t = mm.t;
    const float lr = mm.lr, wd = mm.wd;
    const bool skip = mm.skip;
    do {
      if (mode == kAllreduce) {
        if (!skip) {
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] *= h_rescale;
          fan(e, g);
        }
        break;
      }
      float w[8];
      if (!skip) {
        if (a.w32 != nullptr) ld8<float>(a.w32 + (e - so), w);
        else ld8<WT>(reinterpret_cast<const WT*>(a.weights[a.rank]) + e, w);
      }
      if (mode == kSgd) {
        if (skip) break;
        SgdHyper h{h_lr, h_wd, h_mom, h_rescale, h_clip, a.nesterov, a.has_mom};
        float m[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (a.has_mom) ld8<float>(a.mom + (e - so), m);
#pragma unroll
        for (int j = 0; j < 8; ++j) sgd_step(w[j], g[j], m[j], h, lr, wd);
        if (a.has_mom) st8<float>(a.mom + (e - so), m);
      } else if (mode == kAdam) {
        if (skip) break;
        AdamHyper h{h_lr, h_wd, a.beta1, a.beta2, a.eps, h_rescale, h_clip};
        float m[8], vv[8];
        ld8<float>(a.mom + (e - so), m);
        ld8<float>(a.var + (e - so), vv);
#pragma unroll
        for (int j = 0; j < 8; ++j) adam_step(w[j], g[j], m[j], vv[j], h, lr, wd);
        st8<float>(a.mom + (e - so), m);
        st8<float>(a.var + (e - so), vv);
      } else if (mode == kLambPhase1) {
        // block-uniform chunk: 512 threads x 8 = 4 chunks per block iteration -> reduce per 128-thread group
        float sw = 0.f, sr = 0.f;
        if (!skip) {
          LambHyper h{a.beta1, a.beta2, a.eps, h_rescale, h_clip, h_bc1, h_bc2};
          float m[8], vv[8], r[8];
          ld8<float>(a.mom + (e - so), m);
          ld8<float>(a.var + (e - so), vv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            r[j] = lamb_dir(w[j], g[j], m[j], vv[j], h, wd);
            sw += w[j] * w[j];
            sr += r[j] * r[j];
          }
          st8<float>(a.mom + (e - so), m);
          st8<float>(a.var + (e - so), vv);
          st8<float>(a.rbuf + (e - so), r);
        }
        // 128 consecutive threads (4 warps) share one chunk => one tensor id
        sw = warp_sum(sw);
        sr = warp_sum(sr);
        if ((threadIdx.x & 31) == 0) {
          red[0][threadIdx.x >> 5] = sw;
          red[1][threadIdx.x >> 5] = sr;
        }
        __syncthreads();
        if ((threadIdx.x & 127) == 0 && !skip) {
          const int wq = threadIdx.x >> 5;
          const float aw = red[0][wq] + red[0][wq + 1] + red[0][wq + 2] + red[0][wq + 3];
          const float arr = red[1][wq] + red[1][wq + 1] + red[1][wq + 2] + red[1][wq + 3];
          atomicAdd(wn2 + t, aw);
          atomicAdd(rn2 + t, arr);
        }
        __syncthreads();
        break;
      } else if (mode == kLambPhase2) {
        if (skip) break;
        float wn = sqrtf(__ldcg(wn2 + t)), rn = sqrtf(__ldcg(rn2 + t));
        if (a.lamb_lower >= 0.f) wn = fmaxf(wn, a.lamb_lower);
        if (a.lamb_upper >= 0.f) wn = fminf(wn, a.lamb_upper);
        const float ratio = (wn > 0.f && rn > 0.f) ? wn / rn : 1.f;
        float r[8];
        ld8<float>(a.rbuf + (e - so), r);
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] -= lr * ratio * r[j];
      } else if (mode == kBroadcast) {
        if (skip) break;
      }
      // write back master + fan the low-precision weights out to every replica
      if (a.w32 != nullptr) st8<float>(a.w32 + (e - so), w);
      fan(e, w);
        } while (0);
  };

  // Gradient of one vector through the selected data path.
  auto gather = [&](const Meta& mm, float (&g)[8]) {
    const int64_t e = mm.e;
    if (ll) {
      // 8 elements = 4 packet pairs per source; spin until every packet carries this launch's flag; sum in RANK order so
      // that all replicas compute bit-identical results
      const uint2* mine = a.ll_recv[a.rank] + ll_base;
      bool ok = true;
#pragma unroll
      for (int q = 0; q < (NR > 0 ? NR : kMaxRanks); ++q) {
        if (q < nr) {
          const uint4* src = reinterpret_cast<const uint4*>(mine + (int64_t)q * a.n + e);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 pk;
            uint32_t spins = 0;
            do {
              asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(pk.x), "=r"(pk.y), "=r"(pk.z), "=r"(pk.w) : "l"(src + j) : "memory");
              if (++spins > kSpinLimit) { ok = false; break; }
            } while (pk.y != ll_flag || pk.w != ll_flag);
            g[2 * j] += __uint_as_float(pk.x);
            g[2 * j + 1] += __uint_as_float(pk.z);
          }
        }
      }
      if (!ok && a.err) atomicExch(a.err, 1 + a.rank);
    } else if (mc_grads) {
      // the switch sums the N copies: one load per 16 bytes instead of N
      mc_ld_reduce8<GT>(reinterpret_cast<const GT*>(a.grads_mc) + e, g);
    } else {
      // own copy first, then peers in a rank-rotated order to spread link load
#pragma unroll
      for (int k = 0; k < (NR > 0 ? NR : kMaxRanks); ++k) {
        if (k < nr) {
          const int p = (a.rank + k) % nr;
          float x[8];
          ld8<GT>(reinterpret_cast<const GT*>(a.grads[p]) + e, x);
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] += x[j];
        }
      }
    }
  };

  // One sweep over the owned slice.  Block-uniform trip count (LAMB phase 1 has __syncthreads inside); base is a multiple of
  // 128 vectors, so every group of 128 consecutive threads works on exactly one 1024-element chunk (= one tensor).
  // Bandwidth-bound modes process kU vectors per thread and iteration with all their remote loads issued before the first
  // use: a peer load takes ~2-3 us over NVLink, one vector per iteration left the links latency-bound (0.29 of the link
  // roofline at 1 GB, profiles/bandwidth_sweep_n8_r2_call57.json).
  constexpr int kU = 4;
  auto sweep = [&](const int mode, float* wn2, float* rn2) {
    const bool needs_grad = mode != kLambPhase2 && mode != kBroadcast;
    const int64_t G = (int64_t)gridDim.x * kThreads;
    int64_t base = v_begin + (int64_t)blockIdx.x * kThreads;
    if (!ll && a.pad0 != 0 && (mode == kAllreduce || mode == kSgd || mode == kAdam)) {
      for (; base < v_end; base += kU * G) {
        Meta mm[kU];
        float g[kU][8];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          mm[u] = meta_of(base + u * G + threadIdx.x);
#pragma unroll
          for (int j = 0; j < 8; ++j) g[u][j] = 0.f;
        }
        if (mc_grads) {
#pragma unroll
          for (int u = 0; u < kU; ++u)
            if (!mm[u].skip) mc_ld_reduce8<GT>(reinterpret_cast<const GT*>(a.grads_mc) + mm[u].e, g[u]);
        } else {
#pragma unroll
          for (int k = 0; k < (NR > 0 ? NR : kMaxRanks); ++k) {
            if (k < nr) {
              const GT* src = reinterpret_cast<const GT*>(a.grads[(a.rank + k) % nr]);
              float x[kU][8];
#pragma unroll
              for (int u = 0; u < kU; ++u) {
                if (!mm[u].skip) ld8<GT>(src + mm[u].e, x[u]);
                else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) x[u][j] = 0.f;
                }
              }
#pragma unroll
              for (int u = 0; u < kU; ++u)
#pragma unroll
                for (int j = 0; j < 8; ++j) g[u][j] += x[u][j];
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) process(mode, mm[u], g[u], wn2, rn2);
      }
      return;
    }
    for (; base < v_end; base += G) {
      const Meta mm = meta_of(base + threadIdx.x);
      float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (!mm.skip && needs_grad) gather(mm, g);
      process(mode, mm, g, wn2, rn2);
    }
  };

  if (a.mode == kLamb) {
    // LAMB in ONE launch: phase 1 (reduce, moments, direction, partial norms of my slice) -> all my CTAs done ->
    // cross-rank sum of the per-tensor partial norms over peer memory -> phase 2 (trust ratio, update, fan-out).
    // No NCCL call, no second launch.
    float* my_w = a.norms[a.rank];
    float* my_r = my_w + a.nt_pad;
    sweep(kLambPhase1, my_w, my_r);
    grid_barrier(a.gsync);
    float *tot_w = my_w, *tot_r = my_r;
    if (fan_out) {
      if (!peer_barrier(a, 4u * h_seq + 1u)) return;           // every rank's partial norms are complete and visible
      tot_w = a.norm_tot;
      tot_r = a.norm_tot + a.nt_pad;
      for (int i = blockIdx.x * kThreads + threadIdx.x; i < 2 * a.nt_pad; i += gridDim.x * kThreads) {
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < (NR > 0 ? NR : kMaxRanks); ++k)
          if (k < nr) sum += *(reinterpret_cast<const volatile float*>(a.norms[(a.rank + k) % nr]) + i);
        a.norm_tot[i] = sum;
      }
      grid_barrier(a.gsync);
    }
    sweep(kLambPhase2, tot_w, tot_r);
  } else {
    sweep(a.mode, a.w_norm2, a.r_norm2);
  }
  if (nr > 1 && !ll) {
    if (!peer_barrier(a, 4u * h_seq + 2u)) return;
  }
  if (a.mode == kLamb) {
    // peers have finished reading my partial norms: clear them for the next step
    float* my_w = a.norms[a.rank];
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < 2 * a.nt_pad; i += gridDim.x * kThreads) my_w[i] = 0.f;
  }
  // Everyone has consumed this rank's gradients: clear them so the next backward can accumulate
  // without a separate memset (grad_req='add'-style arena).  Done after the barrier on the FULL local arena.
  if (a.zero_grad && a.mode != kLambPhase2 && a.mode != kBroadcast) {
    const int64_t nv = a.n / 8;
    const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < nv; v += (int64_t)gridDim.x * kThreads)
      st8<GT>(reinterpret_cast<GT*>(a.grads[a.rank]) + v * 8, z);
  }
  if (a.seq_dev != nullptr) {
    // every CTA has read the sequence number long ago (before its first barrier); the last one to finish advances it
    __syncthreads();
    if (threadIdx.x == 0) {
      if (atomicAdd(a.seq_dev + 1, 1u) == gridDim.x - 1) {
        a.seq_dev[1] = 0;
        a.seq_dev[0] = a.seq_dev[0] + 1u;
        if (ll) a.ll_ctr[0] = ll_cnt + 1u;
        __threadfence();
      }
    }
  }
}

template <typename GT, typename WT>
static void launch_nr(const Args& a, int blocks, cudaStream_t s) {
  switch (a.nranks) {
    case 1: fused_kernel<GT, WT, 1><<<blocks, kThreads, 0, s>>>(a); break;
    case 2: fused_kernel<GT, WT, 2><<<blocks, kThreads, 0, s>>>(a); break;
    case 4: fused_kernel<GT, WT, 4><<<blocks, kThreads, 0, s>>>(a); break;
    case 8: fused_kernel<GT, WT, 8><<<blocks, kThreads, 0, s>>>(a); break;
    default: fused_kernel<GT, WT, 0><<<blocks, kThreads, 0, s>>>(a); break;
  }
}

}  // namespace ar
}  // namespace mxb

using namespace mxb;
using namespace mxb::ar;

MXB_API int MXBFusedAllreduceArgsSize() { return (int)sizeof(Args); }

MXB_API int MXBFusedAllreduce(const void* args_blob, int blob_bytes, int blocks, void* stream) {
  MXB_TRY
  MXB_CHECK(blob_bytes == (int)sizeof(Args), "Args struct size mismatch between python and native");
  Args a = *reinterpret_cast<const Args*>(args_blob);
  MXB_CHECK(a.nranks >= 1 && a.nranks <= kMaxRanks, "nranks out of range");
  MXB_CHECK(a.n % kChunk == 0, "arena size must be a multiple of 1024");
  if (blocks <= 0) {
    // one CTA per SM for large arenas (the step runs alone at the end of the iteration: measured 1.5 ms -> see
    // profiles/ for LAMB on 110 M parameters with the previous fixed 64 CTAs); small arenas use fewer CTAs so the
    // flag barrier stays cheap.  All CTAs must be co-resident (they pair up with the same CTA index on the peers).
    const int64_t vec_per_rank = a.n / 8 / (a.nranks > 0 ? a.nranks : 1);
    int64_t want = (vec_per_rank + kThreads * 4 - 1) / (kThreads * 4);
    blocks = (int)(want < 8 ? 8 : (want > kNumSMs ? kNumSMs : want));
  }
  if (blocks > kMaxBlocks) blocks = kMaxBlocks;
  if (a.mode == kLamb) {
    MXB_CHECK(a.gsync != nullptr && a.norm_tot != nullptr && a.norms[a.rank] != nullptr && a.nt_pad > 0,
              "fused LAMB needs the norm buffers and the grid-barrier words");
    if (blocks > kNumSMs) blocks = kNumSMs;     // the intra-rank grid barrier needs every CTA resident
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (a.g_dt == kBF16 && a.w_dt == kBF16) launch_nr<__nv_bfloat16, __nv_bfloat16>(a, blocks, s);
  else if (a.g_dt == kBF16 && a.w_dt == kF32) launch_nr<__nv_bfloat16, float>(a, blocks, s);
  else if (a.g_dt == kF32 && a.w_dt == kF32) launch_nr<float, float>(a, blocks, s);
  else if (a.g_dt == kF32 && a.w_dt == kBF16) launch_nr<float, __nv_bfloat16>(a, blocks, s);
  else throw std::runtime_error("fused allreduce: dtypes must be bf16 or fp32");
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}

// ---- peer-memory plumbing (CommDevice::EnableP2P / InitMergeBuffer analogue, comm.h:678-755) ----
MXB_API int MXBDeviceMalloc(int64_t bytes, void** out) {
  MXB_TRY
  CUDA_CHECK(cudaMalloc(out, (size_t)bytes));
  CUDA_CHECK(cudaMemset(*out, 0, (size_t)bytes));
  MXB_CATCH
}
MXB_API int MXBDeviceFree(void* p) {
  MXB_TRY
  CUDA_CHECK(cudaFree(p));
  MXB_CATCH
}
MXB_API int MXBIpcGetHandle(void* p, void* handle64) {
  MXB_TRY
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  CUDA_CHECK(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), p));
  MXB_CATCH
}
MXB_API int MXBIpcOpenHandle(const void* handle64, void** out) {
  MXB_TRY
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  CUDA_CHECK(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  MXB_CATCH
}
MXB_API int MXBIpcCloseHandle(void* p) {
  MXB_TRY
  CUDA_CHECK(cudaIpcCloseMemHandle(p));
  MXB_CATCH
}
MXB_API int MXBEnablePeerAccess(int dev, int peer) {
  MXB_TRY
  int cur;
  CUDA_CHECK(cudaGetDevice(&cur));
  CUDA_CHECK(cudaSetDevice(dev));
  int can = 0;
  CUDA_CHECK(cudaDeviceCanAccessPeer(&can, dev, peer));
  if (!can) {
    cudaSetDevice(cur);
    throw std::runtime_error("device " + std::to_string(dev) + " cannot access peer " + std::to_string(peer));
  }
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
  } else if (e != cudaSuccess) {
    cudaSetDevice(cur);
    CUDA_CHECK(e);
  }
  CUDA_CHECK(cudaSetDevice(cur));
  MXB_CATCH
}
MXB_API int MXBHostAllocMapped(int64_t bytes, void** host_ptr, void** dev_ptr) {
  MXB_TRY
  CUDA_CHECK(cudaHostAlloc(host_ptr, (size_t)bytes, cudaHostAllocMapped | cudaHostAllocPortable));
  memset(*host_ptr, 0, (size_t)bytes);
  CUDA_CHECK(cudaHostGetDevicePointer(dev_ptr, *host_ptr, 0));
  MXB_CATCH
}
MXB_API int MXBHostFree(void* p) {
  MXB_TRY
  CUDA_CHECK(cudaFreeHost(p));
  MXB_CATCH
}
