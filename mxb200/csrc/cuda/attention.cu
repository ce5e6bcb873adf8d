// Fused multi-head self-attention for sequences of up to 128 tokens (head dim 64, bf16) on tcgen05: scores, masking,
// softmax, dropout and the value product of one (batch, head) never leave the SM.
//
// Reference building blocks this replaces (the reference has no attention operator; a BERT written against it composes
// batch_dot + softmax + Dropout + batch_dot = four kernels and two [B, heads, S, S] round trips through HBM:
// src/operator/tensor/dot-inl.cuh (batch_dot -> cuBLAS), src/operator/nn/softmax.cc, src/operator/nn/dropout-inl.h;
// src/operator/contrib/transformer.cu only holds a scaling helper).
//
// One CTA per (batch, head).  Layout of q / k / v / out: [B * S, H] row-major matrices, head h in columns [64h, 64h+64)
// -- exactly what the three projection GEMMs write, so no head transposes exist anywhere.
//
//   forward   S = Q K^T            tcgen05.mma 128 x 128 x 64     (Q, K: K-major SW128 tiles loaded by TMA)
//             P = softmax(scale * S + key mask) . dropout          registers: each of 128 threads owns one query row
//             O = P V               tcgen05.mma 128 x 64 x 128    (P: bf16 in smem, K-major; V: MN-major)
//   backward  S, P recomputed from the saved log-sum-exp; dV = P~^T dO; dP = dO V^T; dS = P * (dP~ - D) * scale;
//             dQ = dS K; dK = dS^T Q      five more tcgen05 GEMMs, P~ / dS staged once in smem and used both as a K-major
//             and as an MN-major operand (same bytes, different descriptor)
//
// Accumulators live in TMEM (S / dP: 128 columns each, O / dV / dQ / dK: 64 columns each).  Dropout uses Philox4x32-10
// keyed by (seed, offset) with the element index as counter, so the backward pass regenerates the mask instead of storing it.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace mxb {
namespace attn {

constexpr int kS = 128;          // queries / keys per CTA (sequence length <= 128)
constexpr int kD = 64;           // head dimension
constexpr int kThreads = 160;    // warps 0-3: one query row per thread; warp 4: TMA + MMA issuer
constexpr int kTile = kS * kD * 2;   // 16 KB: one [128 x 64] bf16 tile, 128B-swizzled

struct Params {
  int B, S, nh;
  int64_t ld;                    // row stride (elements) of q / k / v / out / dout matrices
  float scale;                   // 1 / sqrt(head dim)
  const uint8_t* key_mask;       // [B, S] 1 = attend, 0 = padding; or null
  float p_drop;
  uint64_t seed, offset;         // Philox key / stream: effective offset = offset + *offset_dev (graph replays advance it)
  const uint64_t* offset_dev;    // device-resident step counter or null
  uint64_t* rng_state;           // forward: [seed, effective offset] written here; backward: read from here
  __nv_bfloat16* out;            // forward
  float* lse;                    // [B, nh, S] log-sum-exp of the scaled, masked scores (natural log)
  const __nv_bfloat16* o_in;     // backward: forward output
  const __nv_bfloat16* dout;
  __nv_bfloat16 *dq, *dk, *dv;
};

// ---------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ uint4 philox(uint64_t seed, uint64_t ctr_hi, uint64_t ctr_lo) {
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Row `row` of a [128 x 128] bf16 matrix stored as two 128B-swizzled 64-column panels (16 KB each): write 8 values at
// column 8 * chunk.
__device__ __forceinline__ void store_row_chunk(uint32_t base, int row, int chunk, const float* v) {
  const uint32_t panel = base + (uint32_t)(chunk >> 3) * kTile;
  const uint32_t addr = panel + (uint32_t)row * 128u + ((((uint32_t)chunk & 7u) ^ ((uint32_t)row & 7u)) << 4);
  sts16(addr, pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}

// Descriptor constants (see gemm_tcgen05.cu): K-major tile = rows of 128 B, 8-row groups 1024 B apart;
// MN-major tile = [64 k][64 mn] boxes of 8 KB, 8-deep k groups 1024 B apart, 64-element mn chunks LBO apart.
constexpr uint32_t kSBO = 1024;

__device__ __forceinline__ void mma_kmajor_a_kmajor_b(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, int n, int ksteps,
                                                      uint32_t a_panel, uint32_t b_panel, bool accumulate, bool issue) {
  // K advances 16 elements = 32 B inside a 64-wide panel; panels are `*_panel` bytes apart
  const uint32_t idesc = ptx::make_idesc_f16(128, n, true, false, false);
  for (int k = 0; k < ksteps; ++k) {
    const uint32_t ao = (uint32_t)(k >> 2) * a_panel + (uint32_t)(k & 3) * 32u;
    const uint32_t bo = (uint32_t)(k >> 2) * b_panel + (uint32_t)(k & 3) * 32u;
    if (issue) ptx::umma_f16(tmem_d, ptx::make_smem_desc(a_addr + ao, 16, kSBO), ptx::make_smem_desc(b_addr + bo, 16, kSBO), idesc,
                  (accumulate || k > 0) ? 1u : 0u);
  }
}
// A K-major (panels), B MN-major ([k rows][n cols], n <= 64 per chunk; chunks of 64 n are `b_chunk` bytes apart)
__device__ __forceinline__ void mma_kmajor_a_mnmajor_b(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, int n, int ksteps,
                                                       uint32_t a_panel, uint32_t b_chunk, bool accumulate, bool issue) {
  const uint32_t idesc = ptx::make_idesc_f16(128, n, true, false, true);
  for (int k = 0; k < ksteps; ++k) {
    const uint32_t ao = (uint32_t)(k >> 2) * a_panel + (uint32_t)(k & 3) * 32u;
    const uint32_t bo = (uint32_t)k * (16u * 128u);          // 16 k rows of 128 B
    if (issue) ptx::umma_f16(tmem_d, ptx::make_smem_desc(a_addr + ao, 16, kSBO), ptx::make_smem_desc(b_addr + bo, b_chunk, kSBO), idesc,
                  (accumulate || k > 0) ? 1u : 0u);
  }
}
// A MN-major ([k rows][m cols]: two 64-wide m chunks `a_chunk` bytes apart), B MN-major
__device__ __forceinline__ void mma_mnmajor_a_mnmajor_b(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, int n, int ksteps,
                                                        uint32_t a_chunk, uint32_t b_chunk, bool accumulate, bool issue) {
  const uint32_t idesc = ptx::make_idesc_f16(128, n, true, true, true);
  for (int k = 0; k < ksteps; ++k) {
    const uint32_t o = (uint32_t)k * (16u * 128u);
    if (issue) ptx::umma_f16(tmem_d, ptx::make_smem_desc(a_addr + o, a_chunk, kSBO), ptx::make_smem_desc(b_addr + o, b_chunk, kSBO), idesc,
                  (accumulate || k > 0) ? 1u : 0u);
  }
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  ptx::tmem_ld_32x32b_x32(taddr, r);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

struct Smem {
  uint8_t* q;
  uint8_t* k;
  uint8_t* v;
  uint8_t* dout;     // backward only
  uint8_t* p;        // [128 x 128] bf16, two panels (32 KB)
  uint8_t* ds;       // backward only
  uint64_t* bars;
  uint32_t* tmem_slot;
};

__device__ __forceinline__ float key_bias(const Params& p, int b, int key) {
  if (key >= p.S) return -INFINITY;
  if (p.key_mask != nullptr && p.key_mask[(int64_t)b * p.S + key] == 0) return -INFINITY;
  return 0.f;
}

// ================================================================================================ forward
__global__ void __launch_bounds__(kThreads, 3)
fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
           const __grid_constant__ CUtensorMap tmV, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sq = sm;
  uint8_t* sk = sm + kTile;
  uint8_t* sv = sm + 2 * kTile;
  uint8_t* sp = sm;                             // P (2 panels) overwrites Q | K once S = Q K^T has completed
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 3 * kTile);   // 0: q,k,v landed  1: S ready  2: P staged  3: O ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int b = blockIdx.x / p.nh, h = blockIdx.x % p.nh;
  const uint64_t rng_off = p.offset + (p.offset_dev != nullptr ? *p.offset_dev : 0ull);
  if (blockIdx.x == 0 && threadIdx.x == 0 && p.rng_state != nullptr) {
    p.rng_state[0] = p.seed;
    p.rng_state[1] = rng_off;
  }

  if (warp == 4 && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::mbar_init(&bars[2], 1);
    ptx::mbar_init(&bars[3], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_slot, 128);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_s = tmem, t_o = tmem;        // O reuses the first 64 score columns (every row was drained before bars[2])

  if (warp == 4) {
    {   // the whole warp walks this code (uniform registers for descriptors); the elected lane issues
      const bool issuer = ptx::elect_one();
      if (issuer) ptx::mbar_arrive_expect_tx(&bars[0], 3 * kTile);
      if (issuer) ptx::tma_load_2d(sq, &tmQ, &bars[0], h * kD, b * p.S);
      if (issuer) ptx::tma_load_2d(sk, &tmK, &bars[0], h * kD, b * p.S);
      if (issuer) ptx::tma_load_2d(sv, &tmV, &bars[0], h * kD, b * p.S);
      ptx::mbar_wait(&bars[0], 0);
      ptx::tc_fence_after();
      mma_kmajor_a_kmajor_b(t_s, ptx::smem_u32(sq), ptx::smem_u32(sk), 128, 4, kTile, kTile, false, issuer);
      if (issuer) ptx::umma_commit(&bars[1]);
      ptx::mbar_wait(&bars[2], 0);
      ptx::tc_fence_after();
      // O[q, d] = sum_key P[q, key] V[key, d]: P K-major (two 64-key panels), V MN-major (two [64 key][64 d] boxes)
      mma_kmajor_a_mnmajor_b(t_o, ptx::smem_u32(sp), ptx::smem_u32(sv), 64, 8, kTile, 64 * 128, false, issuer);
      if (issuer) ptx::umma_commit(&bars[3]);
    }
  } else {
    const int row = warp * 32 + lane;                      // query owned by this thread (TMEM lane = row)
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    ptx::mbar_wait(&bars[1], 0);
    ptx::tc_fence_after();
    const float sl2 = p.scale * 1.4426950408889634f;       // scores are scaled and moved to the exp2 domain at once
    // pass 1: row maximum
    float mx = -INFINITY;
    for (int c = 0; c < 4; ++c) {
      float v[32];
      tmem_ld32(t_s + lane_base + c * 32, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, v[j] * sl2 + key_bias(p, b, c * 32 + j));
    }
    if (mx == -INFINITY) mx = 0.f;                         // fully masked row: all probabilities zero
    // pass 2: probabilities (un-normalised), dropout, staging as the A operand of the PV product
    const float keep = 1.f - p.p_drop;
    const float inv_keep = p.p_drop > 0.f ? 1.f / keep : 1.f;
    const uint32_t thresh = (uint32_t)fminf(4294967295.f, keep * 4294967296.f);
    float sum = 0.f;
    const uint32_t sp_addr = ptx::smem_u32(sp);
    const uint64_t row_ctr = ((uint64_t)blockIdx.x * kS + (uint64_t)row) * (kS / 4);
    for (int c = 0; c < 4; ++c) {
      float v[32];
      tmem_ld32(t_s + lane_base + c * 32, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float e = exp2f(v[j] * sl2 + key_bias(p, b, c * 32 + j) - mx);
        sum += e;
        v[j] = e;
      }
      if (p.p_drop > 0.f) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const uint4 r = philox(p.seed, rng_off, row_ctr + (uint64_t)(c * 8 + j4));
          v[j4 * 4 + 0] = r.x <= thresh ? v[j4 * 4 + 0] * inv_keep : 0.f;
          v[j4 * 4 + 1] = r.y <= thresh ? v[j4 * 4 + 1] * inv_keep : 0.f;
          v[j4 * 4 + 2] = r.z <= thresh ? v[j4 * 4 + 2] * inv_keep : 0.f;
          v[j4 * 4 + 3] = r.w <= thresh ? v[j4 * 4 + 3] * inv_keep : 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) store_row_chunk(sp_addr, row, c * 4 + i, v + i * 8);
    }
    ptx::tc_fence_before();
    ptx::fence_proxy_async_smem();
    bar_sync(1, 128);
    if (threadIdx.x == 0) ptx::mbar_arrive(&bars[2]);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    if (row < p.S && p.lse != nullptr)
      p.lse[((int64_t)b * p.nh + h) * p.S + row] = sum > 0.f ? (mx * 0.6931471805599453f + logf(sum)) : -INFINITY;
    ptx::mbar_wait(&bars[3], 0);
    ptx::tc_fence_after();
    __nv_bfloat16* dst = p.out + ((int64_t)b * p.S + row) * p.ld + h * kD;
    for (int c = 0; c < 2; ++c) {
      float v[32];
      tmem_ld32(t_o + lane_base + c * 32, v);
      if (row < p.S) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(dst + c * 32 + i * 8) =
              make_uint4(pack2(v[i * 8 + 0] * inv, v[i * 8 + 1] * inv), pack2(v[i * 8 + 2] * inv, v[i * 8 + 3] * inv),
                         pack2(v[i * 8 + 4] * inv, v[i * 8 + 5] * inv), pack2(v[i * 8 + 6] * inv, v[i * 8 + 7] * inv));
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 128);
  }
}

// ================================================================================================ backward
__global__ void __launch_bounds__(kThreads, 2)
bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
           const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sq = sm;
  uint8_t* sk = sm + kTile;
  uint8_t* sdo = sm + 2 * kTile;
  uint8_t* sv = sm + 3 * kTile;
  uint8_t* sp = sm + 3 * kTile;                 // P~  (2 panels): panel 0 overwrites V once dP = dO V^T has completed
  uint8_t* sds = sm + 5 * kTile;                // dS  (2 panels)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 7 * kTile);
  // 0: tiles landed  1: S + dP ready  2: P~ and dS staged  3: dV, dQ, dK ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int b = blockIdx.x / p.nh, h = blockIdx.x % p.nh;

  if (warp == 4 && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    ptx::prefetch_tmap(&tmDO);
    for (int i = 0; i < 4; ++i) ptx::mbar_init(&bars[i], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // dV / dQ / dK reuse the S and dP columns: all 128 rows have been drained into P~ / dS before bars[2] is signalled
  const uint32_t t_s = tmem, t_dp = tmem + 128, t_dv = tmem, t_dq = tmem + 64, t_dk = tmem + 128;

  if (warp == 4) {
    {   // the whole warp walks this code (uniform registers for descriptors); the elected lane issues
      const bool issuer = ptx::elect_one();
      if (issuer) ptx::mbar_arrive_expect_tx(&bars[0], 4 * kTile);
      if (issuer) ptx::tma_load_2d(sq, &tmQ, &bars[0], h * kD, b * p.S);
      if (issuer) ptx::tma_load_2d(sk, &tmK, &bars[0], h * kD, b * p.S);
      if (issuer) ptx::tma_load_2d(sv, &tmV, &bars[0], h * kD, b * p.S);
      if (issuer) ptx::tma_load_2d(sdo, &tmDO, &bars[0], h * kD, b * p.S);
      ptx::mbar_wait(&bars[0], 0);
      ptx::tc_fence_after();
      // S = Q K^T and dP = dO V^T: both [query x key], K = head dim, all operands K-major
      mma_kmajor_a_kmajor_b(t_s, ptx::smem_u32(sq), ptx::smem_u32(sk), 128, 4, kTile, kTile, false, issuer);
      mma_kmajor_a_kmajor_b(t_dp, ptx::smem_u32(sdo), ptx::smem_u32(sv), 128, 4, kTile, kTile, false, issuer);
      if (issuer) ptx::umma_commit(&bars[1]);
      ptx::mbar_wait(&bars[2], 0);
      ptx::tc_fence_after();
      // dV[key, d] = sum_q P~[q, key] dO[q, d] : A = P~ read MN-major (m = key), B = dO read MN-major (n = d); k = query
      mma_mnmajor_a_mnmajor_b(t_dv, ptx::smem_u32(sp), ptx::smem_u32(sdo), 64, 8, kTile, 64 * 128, false, issuer);
      // dQ[q, d] = sum_key dS[q, key] K[key, d] : A = dS K-major (two key panels), B = K MN-major
      mma_kmajor_a_mnmajor_b(t_dq, ptx::smem_u32(sds), ptx::smem_u32(sk), 64, 8, kTile, 64 * 128, false, issuer);
      // dK[key, d] = sum_q dS[q, key] Q[q, d] : A = dS MN-major, B = Q MN-major
      mma_mnmajor_a_mnmajor_b(t_dk, ptx::smem_u32(sds), ptx::smem_u32(sq), 64, 8, kTile, 64 * 128, false, issuer);
      if (issuer) ptx::umma_commit(&bars[3]);
    }
  } else {
    const int row = warp * 32 + lane;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    // D[q] = sum_d dO[q, d] * O[q, d]  (straight from global memory: 2 x 128 B per thread)
    float dsum = 0.f;
    if (row < p.S) {
      const __nv_bfloat16* po = p.o_in + ((int64_t)b * p.S + row) * p.ld + h * kD;
      const __nv_bfloat16* pd = p.dout + ((int64_t)b * p.S + row) * p.ld + h * kD;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 uo = *reinterpret_cast<const uint4*>(po + i * 8), ud = *reinterpret_cast<const uint4*>(pd + i * 8);
        const __nv_bfloat162* ho = reinterpret_cast<const __nv_bfloat162*>(&uo);
        const __nv_bfloat162* hd = reinterpret_cast<const __nv_bfloat162*>(&ud);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dsum += __low2float(ho[j]) * __low2float(hd[j]) + __high2float(ho[j]) * __high2float(hd[j]);
      }
    }
    const float lse = (row < p.S) ? p.lse[((int64_t)b * p.nh + h) * p.S + row] : -INFINITY;
    const float sl2 = p.scale * 1.4426950408889634f;
    const float lse2 = lse * 1.4426950408889634f;
    const float keep = 1.f - p.p_drop;
    const float inv_keep = p.p_drop > 0.f ? 1.f / keep : 1.f;
    const uint32_t thresh = (uint32_t)fminf(4294967295.f, keep * 4294967296.f);
    const uint64_t row_ctr = ((uint64_t)blockIdx.x * kS + (uint64_t)row) * (kS / 4);
    const uint64_t rng_seed = p.rng_state != nullptr ? p.rng_state[0] : p.seed;
    const uint64_t rng_off = p.rng_state != nullptr ? p.rng_state[1] : p.offset;
    ptx::mbar_wait(&bars[1], 0);
    ptx::tc_fence_after();
    const uint32_t sp_addr = ptx::smem_u32(sp), sds_addr = ptx::smem_u32(sds);
    for (int c = 0; c < 4; ++c) {
      float s[32], dp[32];
      tmem_ld32(t_s + lane_base + c * 32, s);
      tmem_ld32(t_dp + lane_base + c * 32, dp);
      uint32_t keepbits = 0xFFFFFFFFu;
      if (p.p_drop > 0.f) {
        keepbits = 0;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const uint4 r = philox(rng_seed, rng_off, row_ctr + (uint64_t)(c * 8 + j4));
          keepbits |= (r.x <= thresh ? 1u : 0u) << (j4 * 4 + 0);
          keepbits |= (r.y <= thresh ? 1u : 0u) << (j4 * 4 + 1);
          keepbits |= (r.z <= thresh ? 1u : 0u) << (j4 * 4 + 2);
          keepbits |= (r.w <= thresh ? 1u : 0u) << (j4 * 4 + 3);
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float bias = key_bias(p, b, c * 32 + j);
        float pj = (lse == -INFINITY || bias == -INFINITY) ? 0.f : exp2f(s[j] * sl2 - lse2);   // normalised probability
        const bool kept = (keepbits >> j) & 1u;
        const float pt = kept ? pj * inv_keep : 0.f;                 // P~ = dropout(P): operand of dV
        const float dpj = kept ? dp[j] * inv_keep : 0.f;             // gradient w.r.t. P
        s[j] = pt;
        dp[j] = pj * (dpj - dsum) * p.scale;                         // dS (w.r.t. the unscaled scores)
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        store_row_chunk(sp_addr, row, c * 4 + i, s + i * 8);
        store_row_chunk(sds_addr, row, c * 4 + i, dp + i * 8);
      }
    }
    ptx::tc_fence_before();
    ptx::fence_proxy_async_smem();
    bar_sync(1, 128);
    if (threadIdx.x == 0) ptx::mbar_arrive(&bars[2]);
    ptx::mbar_wait(&bars[3], 0);
    ptx::tc_fence_after();
    // row = key index for dV / dK, query index for dQ
    const int64_t goff = ((int64_t)b * p.S + row) * p.ld + h * kD;
    auto drain = [&](uint32_t taddr, __nv_bfloat16* base) {
      for (int c = 0; c < 2; ++c) {
        float v[32];
        tmem_ld32(taddr + lane_base + c * 32, v);
        if (row < p.S) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(base + goff + c * 32 + i * 8) =
                make_uint4(pack2(v[i * 8 + 0], v[i * 8 + 1]), pack2(v[i * 8 + 2], v[i * 8 + 3]),
                           pack2(v[i * 8 + 4], v[i * 8 + 5]), pack2(v[i * 8 + 6], v[i * 8 + 7]));
        }
      }
    };
    drain(t_dv, p.dv);
    drain(t_dq, p.dq);
    drain(t_dk, p.dk);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 256);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  if (!fn) throw std::runtime_error("cuTensorMapEncodeTiled not available from the driver");
  return fn;
}

static CUtensorMap map_rows(const void* ptr, int64_t rows, int64_t cols, int64_t ld) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kD, (cuuint32_t)kS};
  cuuint32_t estr[2] = {1, 1};
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (strides[0] & 15) != 0)
    throw std::runtime_error("attention: operands must be 16-byte aligned with a row stride multiple of 8");
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("attention: cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return m;
}

}  // namespace attn
}  // namespace mxb

using namespace mxb;

// q, k, v, out: [B*S, ld] bf16 matrices (head h in columns 64h .. 64h+63); lse: fp32 [B, nh, S]; key_mask: uint8 [B, S] or null
MXB_API int MXBAttnFwd(const void* q, const void* k, const void* v, void* out, float* lse, int B, int S, int nh, int64_t ld,
                       float scale, const void* key_mask, float p_drop, uint64_t seed, uint64_t offset, const void* offset_dev,
                       void* rng_state, void* stream) {
  MXB_TRY
  MXB_CHECK(S >= 1 && S <= attn::kS && B >= 1 && nh >= 1, "attention kernel: sequence length must be in [1, 128]");
  MXB_CHECK(ld >= (int64_t)nh * attn::kD, "attention kernel: head dimension is 64");
  attn::Params p = {};
  p.B = B; p.S = S; p.nh = nh; p.ld = ld; p.scale = scale;
  p.key_mask = reinterpret_cast<const uint8_t*>(key_mask);
  p.p_drop = p_drop; p.seed = seed; p.offset = offset;
  p.offset_dev = reinterpret_cast<const uint64_t*>(offset_dev);
  p.rng_state = reinterpret_cast<uint64_t*>(rng_state);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  const int64_t rows = (int64_t)B * S, cols = (int64_t)nh * attn::kD;
  CUtensorMap tq = attn::map_rows(q, rows, cols, ld), tk = attn::map_rows(k, rows, cols, ld), tv = attn::map_rows(v, rows, cols, ld);
  const int smem = 3 * attn::kTile + 256 + 1024;
  CUDA_CHECK(cudaFuncSetAttribute(attn::fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  attn::fwd_kernel<<<B * nh, attn::kThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tq, tk, tv, p);
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}

MXB_API int MXBAttnBwd(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                       void* dq, void* dk, void* dv, int B, int S, int nh, int64_t ld, float scale, const void* key_mask,
                       float p_drop, const void* rng_state, void* stream) {
  MXB_TRY
  MXB_CHECK(S >= 1 && S <= attn::kS && B >= 1 && nh >= 1, "attention kernel: sequence length must be in [1, 128]");
  attn::Params p = {};
  p.rng_state = reinterpret_cast<uint64_t*>(const_cast<void*>(rng_state));
  p.B = B; p.S = S; p.nh = nh; p.ld = ld; p.scale = scale;
  p.key_mask = reinterpret_cast<const uint8_t*>(key_mask);
  p.p_drop = p_drop;
  p.lse = const_cast<float*>(lse);
  p.o_in = reinterpret_cast<const __nv_bfloat16*>(out);
  p.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq);
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk);
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  const int64_t rows = (int64_t)B * S, cols = (int64_t)nh * attn::kD;
  CUtensorMap tq = attn::map_rows(q, rows, cols, ld), tk = attn::map_rows(k, rows, cols, ld), tv = attn::map_rows(v, rows, cols, ld),
              tdo = attn::map_rows(dout, rows, cols, ld);
  const int smem = 7 * attn::kTile + 256 + 1024;
  CUDA_CHECK(cudaFuncSetAttribute(attn::bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  attn::bwd_kernel<<<B * nh, attn::kThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tq, tk, tv, tdo, p);
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}
