// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> smem ring -> tcgen05.mma (accumulators in
// TMEM, double buffered) -> tcgen05.ld epilogue (bias / ReLU / per-column BN statistics) -> swizzled smem
// staging -> TMA store (or TMA reduce-add for split-K).
//
// Role in the framework: this is the FullyConnected kernel (reference: src/operator/nn/fully_connected-inl.h
// -> linalg_gemm -> cublasSgemmEx/GemmEx, linalg_impl.h:192-291) and, in NHWC, every 1x1 convolution of the
// ResNet bottlenecks (fwd, dgrad, wgrad are all plain GEMMs).  Operand "major-ness" is a template parameter
// so that the three backward/forward products run without any transposes in global memory:
//
//   fwd   Y[M,N]  = X[M,K]  . W[N,K]^T      A K-major , B K-major
//   dgrad dX[M,K] = dY[M,N] . W[N,K]        A K-major , B MN-major (W stored [red=N, out=K])
//   wgrad dW[N,K] = dY[M,N]^T . X[M,K]      A MN-major, B MN-major, split along the huge reduction dim M
//
// Tile: BLOCK_M=128 x BLOCK_N in {64,128} x BLOCK_K=64 (one 128-byte swizzle atom of bf16), UMMA 128xBNx16,
// cta_group::1.  Warp roles: warp0 = TMA producer, warp1 = MMA issuer, warp2 = TMEM allocator,
// warps4-11 = epilogue (TMEM lane quarter = warp_idx % 4; the two warps of a quarter split the columns;
// TMEM loads are batched, the staging buffer is double buffered so the TMA store of tile i overlaps the
// epilogue of tile i+1, and per-column BatchNorm statistics are accumulated per CTA in shared memory).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace mxb {
namespace gemm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int kThreads = 384;      // 4 control warps + 2 epilogue groups of 4 warps
constexpr int kGroupThreads = 128; // one epilogue group = one warp per TMEM lane quarter
constexpr int kMaxStatCols = 1024; // per-CTA, per-epilogue-group column-statistics accumulators in shared memory
constexpr int kSlabBytes = BLOCK_M * 128;  // one 128B-swizzled output box: 128 rows x 128 bytes

// n / d for a launch-constant d without an integer divide (tile decode runs once per tile in every role).
struct FastDiv {
  uint32_t d, mul, shift;
  __device__ __forceinline__ uint32_t div(uint32_t n) const {
    return (uint32_t)(((uint64_t)__umulhi(n, mul) + n) >> shift);
  }
};

static FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;
  f.shift = s;
  f.mul = (uint32_t)((((1ull << 32) * ((1ull << s) - d)) / d) + 1);
  return f;
}

struct Params {
  int M, N, K;          // logical problem: D[M,N] = A[M,K] * B[N,K]^T   (K = reduction)
  int num_splits;       // split-K factor (>1 only with REDUCE epilogue)
  int kb_per_split;     // k-blocks per split
  // rasterisation: groups of `gm` m-blocks; inside a group the m-block index runs fastest, so the tiles that share
  // an A panel (same m-block, different n-block) are at most gm tiles apart and hit L2.
  int gm, full_groups, tail_gm;
  FastDiv d_split, d_group, d_gm, d_tail;
  const float* bias;    // [N] or null
  int relu;
  float* col_sum;       // [N] fp32 accumulators (atomicAdd) or null : sum over rows of stored output
  float* col_sumsq;     // [N]
  const __nv_bfloat16* addend;  // optional C[M,N] (row stride ldc): D = A*B + C  (residual-gradient fusion)
  int64_t ldc;
  // debug overrides for descriptor fields (0 = use computed); lets one GPU run probe encodings
  int dbg_lbo_a, dbg_sbo_a, dbg_lbo_b, dbg_sbo_b;
};

template <int BN, typename OutT, int CTAS = 1>
struct Smem {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = (BN / CTAS) * BLOCK_K * 2;   // CTA pair: each CTA stages half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSlabCols = 128 / (int)sizeof(OutT);      // 64 bf16 / 32 fp32 columns per staged slab
  static constexpr int kSlabs = BN / kSlabCols;
  static constexpr int kGroupBufs = 2;                           // staging slabs per epilogue group
  static constexpr int kStagingBytes = 2 * kGroupBufs * kSlabBytes;
  static constexpr int kStatBytes = 2 /*groups*/ * 2 /*sum,sumsq*/ * kMaxStatCols * 4;
  static constexpr int kBudget = 227 * 1024 - 1024 /*align slack*/ - 256 /*barriers*/ - kStatBytes - kStagingBytes;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTotal = kStages * kStageBytes + kStagingBytes + kStatBytes + 256 + 1024;
  static_assert(kStages >= 3, "smem budget too small for a 3-stage pipeline");
};

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void named_bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// CTAS == 2: a thread-block cluster of two CTAs (one TPC) owns a 256 x BN tile.  Each CTA stages 128 rows of A and BN/2
// rows of B, the leader (cluster rank 0) issues tcgen05.mma.cta_group::2 (M = 256), the accumulator rows 0-127 live in
// the leader's TMEM and rows 128-255 in the peer's, and each CTA drains / stores its own half.  Per SM this halves the
// B-operand shared-memory traffic per flop (32 KB instead of 48 KB per 128x256x64 of work).
template <int BN, bool A_MN, bool B_MN, typename OutT, bool REDUCE, int CTAS>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmD, const Params p) {
  using S = Smem<BN, OutT, CTAS>;
  constexpr int kBRows = BN / CTAS;                 // B rows staged by this CTA
  constexpr int kTileM = BLOCK_M * CTAS;            // rows of the output tile owned by the cluster
  const uint32_t cta_rank = CTAS == 2 ? ptx::cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  const int cluster_id = (int)blockIdx.x / CTAS, num_clusters = (int)gridDim.x / CTAS;
  constexpr int kStages = S::kStages;
  constexpr int kTmemCols = 2 * BN;  // double-buffered fp32 accumulators (power of two >= 32)
  constexpr int kSlabCols = S::kSlabCols;
  constexpr int kSlabs = S::kSlabs;
  constexpr int kGroupBufs = S::kGroupBufs;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * S::kABytes;
  uint8_t* smem_d0 = smem + kStages * S::kStageBytes;
  float* colacc = reinterpret_cast<float*>(smem_d0 + S::kStagingBytes);   // [group][sum|sumsq][kMaxStatCols]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(colacc) + S::kStatBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (p.M + kTileM - 1) / kTileM;
  const int num_n = (p.N + BN - 1) / BN;
  const int num_kb_total = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int num_work = num_m * num_n * p.num_splits;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmD);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], CTAS * kSlabs * 4);   // every slab of a tile is drained by the 4 warps of one group (per CTA)
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CTAS == 2) {
      ptx::tmem_alloc_2sm(tmem_base_slot, kTmemCols);   // collective over the pair: same warp id in both CTAs
      ptx::tmem_relinquish_2sm();
    } else {
      ptx::tmem_alloc(tmem_base_slot, kTmemCols);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CTAS == 2) ptx::cluster_sync();   // the peer's barriers must be initialised before anything signals them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  auto decode = [&](int w, int& m_blk, int& n_blk, int& kb0, int& kb1) {
    const uint32_t tile = p.d_split.div((uint32_t)w);
    const int split = w - (int)tile * p.num_splits;
    const uint32_t grp = p.d_group.div(tile);
    const uint32_t within = tile - grp * p.d_group.d;
    uint32_t nb, gm_eff;
    if ((int)grp >= p.full_groups) {
      nb = p.d_tail.div(within);
      gm_eff = p.tail_gm;
    } else {
      nb = p.d_gm.div(within);
      gm_eff = p.gm;
    }
    n_blk = (int)nb;
    m_blk = (int)(grp * p.gm + (within - nb * gm_eff));
    kb0 = split * p.kb_per_split;
    kb1 = kb0 + p.kb_per_split;
    if (kb1 > num_kb_total) kb1 = num_kb_total;
  };

  if (warp == 0) {
    // ===================== TMA producer (one thread) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto load = [&](void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
        if constexpr (CTAS == 2) ptx::tma_load_2d_2sm(dst, tm, bar, c0, c1);   // bytes credited to the leader's barrier
        else ptx::tma_load_2d(dst, tm, bar, c0, c1);
      };
      for (int w = cluster_id; w < num_work; w += num_clusters) {
        int m_blk, n_blk, kb0, kb1;
        decode(w, m_blk, n_blk, kb0, kb1);
        const int m0 = m_blk * kTileM + (int)cta_rank * BLOCK_M;     // first A / output row of this CTA
        const int n0 = n_blk * BN + (int)cta_rank * kBRows;          // first B row staged by this CTA
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (is_leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], CTAS * S::kStageBytes);
          uint8_t* a_dst = smem_a + stage * S::kABytes;
          uint8_t* b_dst = smem_b + stage * S::kBBytes;
          if constexpr (A_MN) {
            // A stored [K, M]: boxes of [64 k][64 m]; chunk c covers m in [64c, 64c+64)
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)
              load(a_dst + c * (BLOCK_K * 128), &tmA, &full_bar[stage], m0 + c * 64, kb * BLOCK_K);
          } else {
            load(a_dst, &tmA, &full_bar[stage], kb * BLOCK_K, m0);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int c = 0; c < kBRows / 64; ++c)
              load(b_dst + c * (BLOCK_K * 128), &tmB, &full_bar[stage], n0 + c * 64, kb * BLOCK_K);
          } else {
            load(b_dst, &tmB, &full_bar[stage], kb * BLOCK_K, n0);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0 && is_leader) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(kTileM, BN, true, A_MN, B_MN);
      // K-major: rows of 128B, 8-row groups 1024B apart (SBO); LBO unused (encoded as 16B).
      // MN-major: 64-element MN chunks LBO apart, 8-deep k groups 1024B apart (SBO).
      const uint32_t lbo_a = p.dbg_lbo_a ? p.dbg_lbo_a : (A_MN ? BLOCK_K * 128 : 16);
      const uint32_t sbo_a = p.dbg_sbo_a ? p.dbg_sbo_a : 1024;
      const uint32_t lbo_b = p.dbg_lbo_b ? p.dbg_lbo_b : (B_MN ? BLOCK_K * 128 : 16);
      const uint32_t sbo_b = p.dbg_sbo_b ? p.dbg_sbo_b : 1024;
      constexpr uint32_t kstep_a = A_MN ? UMMA_K * 128 : UMMA_K * 2;  // bytes per UMMA_K advance
      constexpr uint32_t kstep_b = B_MN ? UMMA_K * 128 : UMMA_K * 2;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = cluster_id; w < num_work; w += num_clusters, ++it) {
        int m_blk, n_blk, kb0, kb1;
        decode(w, m_blk, n_blk, kb0, kb1);
        const int buf = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        ptx::mbar_wait(&tmem_empty[buf], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem_a + stage * S::kABytes);
          const uint32_t b_addr = ptx::smem_u32(smem_b + stage * S::kBBytes);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            uint64_t adesc = ptx::make_smem_desc(a_addr + k * kstep_a, lbo_a, sbo_a);
            uint64_t bdesc = ptx::make_smem_desc(b_addr + k * kstep_b, lbo_b, sbo_b);
            if constexpr (CTAS == 2) ptx::umma_f16_2sm(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else ptx::umma_f16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) when these MMAs retire
          if constexpr (CTAS == 2) ptx::umma_commit_2sm(&empty_bar[stage], 3);
          else ptx::umma_commit(&empty_bar[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        // accumulator ready for the epilogue (of both CTAs)
        if constexpr (CTAS == 2) ptx::umma_commit_2sm(&tmem_full[buf], 3);
        else ptx::umma_commit(&tmem_full[buf]);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: two independent groups of 4 warps =====================
    // A tile's accumulator is drained slab by slab (slab = one 128B-wide swizzled output box: 64 bf16 or 32 fp32
    // columns).  Slabs alternate between the two groups, each group owns kGroupBufs staging slabs and its own
    // named barrier, and one thread per group issues the TMA stores -- so one group's TMEM->smem conversion
    // overlaps the other group's store/statistics, and the per-thread code per slab is ~100 instructions.
    const int g = (warp - 4) >> 2;                // epilogue group
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;                // row of the tile owned by this thread
    const int gtid = (threadIdx.x - 128) & (kGroupThreads - 1);
    const int epi_tid = threadIdx.x - 128;        // 0..255
    const bool leader = (gtid == 0);
    const int bar_id = 1 + g;
    const uint32_t stage_base = ptx::smem_u32(smem_d0 + g * (kGroupBufs * kSlabBytes));
    const uint32_t swz = (uint32_t)(row & 7);
    const bool want_stats = (sizeof(OutT) == 2) && (p.col_sum != nullptr);
    const bool stats_smem = want_stats && (p.N <= kMaxStatCols);
    float* myacc = colacc + g * (2 * kMaxStatCols);   // this group's private accumulators: no atomics needed
    if (stats_smem) {
      for (int i = gtid; i < 2 * kMaxStatCols; i += kGroupThreads) myacc[i] = 0.f;
      named_bar_sync(bar_id, kGroupThreads);
    }
    int it = 0;
    uint32_t slab_seq = 0;   // slabs seen by this CTA (both groups count identically)
    uint32_t my_cnt = 0;     // slabs processed by this group -> staging buffer rotation
    for (int w = cluster_id; w < num_work; w += num_clusters, ++it) {
      int m_blk, n_blk, kb0, kb1;
      decode(w, m_blk, n_blk, kb0, kb1);
      const int buf = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = m_blk * kTileM + (int)cta_rank * BLOCK_M;   // first output row of this CTA
      const int grow = m0 + row;
      bool acc_ready = false;
#pragma unroll 1
      for (int s = 0; s < kSlabs; ++s, ++slab_seq) {
        if ((int)(slab_seq & 1) != g) continue;
        const uint32_t sd = stage_base + (my_cnt % kGroupBufs) * kSlabBytes;
        ++my_cnt;
        const int col0 = n_blk * BN + s * kSlabCols;   // first global column of this slab
        // residual addend: issue the global loads before blocking on the accumulator
        uint4 cadd[sizeof(OutT) == 2 ? 8 : 1];
        if constexpr (sizeof(OutT) == 2) {
          if (p.addend != nullptr) {
            const __nv_bfloat16* src = p.addend + (int64_t)grow * p.ldc + col0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (grow < p.M && col0 + i * 8 < p.N) cadd[i] = *reinterpret_cast<const uint4*>(src + i * 8);
              else cadd[i] = make_uint4(0, 0, 0, 0);
            }
          }
        }
        // the staging slab about to be overwritten must have been drained by its previous TMA store
        if (leader) ptx::tma_wait_group_read<kGroupBufs - 1>();
        named_bar_sync(bar_id, kGroupThreads);
        if (!acc_ready) {
          ptx::mbar_wait(&tmem_full[buf], acc_phase);
          ptx::tc_fence_after();
          acc_ready = true;
        }
        const uint32_t taddr = tmem_base + buf * BN + s * kSlabCols + ((uint32_t)(q * 32) << 16);
        uint32_t r[kSlabCols];
#pragma unroll
        for (int c = 0; c < kSlabCols / 32; ++c) ptx::tmem_ld_32x32b_x32(taddr + c * 32, r + c * 32);
        ptx::tmem_ld_wait();
        // this warp is done with TMEM for this slab: release (the MMA warp waits for all kSlabs*4 arrivals)
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CTAS == 2) ptx::mbar_arrive_leader(&tmem_empty[buf]);   // the MMA issuer lives in the leader
          else ptx::mbar_arrive(&tmem_empty[buf]);
        }
        float v[kSlabCols];
#pragma unroll
        for (int j = 0; j < kSlabCols; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int i = 0; i < kSlabCols / 4; ++i) {
            if (col0 + i * 4 < p.N) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + i);
              v[i * 4 + 0] += b4.x; v[i * 4 + 1] += b4.y; v[i * 4 + 2] += b4.z; v[i * 4 + 3] += b4.w;
            }
          }
        }
        if constexpr (sizeof(OutT) == 2) {
          if (p.addend != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&cadd[i]);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                v[i * 8 + 2 * k] += __low2float(h[k]);
                v[i * 8 + 2 * k + 1] += __high2float(h[k]);
              }
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < kSlabCols; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        const uint32_t rbase = sd + (uint32_t)row * 128u;
        if constexpr (sizeof(OutT) == 2) {
#pragma unroll
          for (int i = 0; i < 8; ++i)   // 8 x 16B chunks = 64 bf16, XOR-swizzled by row (TMA SWIZZLE_128B)
            st_shared_v4(rbase + (((uint32_t)i ^ swz) << 4), pack_bf16(v[i * 8 + 0], v[i * 8 + 1]),
                         pack_bf16(v[i * 8 + 2], v[i * 8 + 3]), pack_bf16(v[i * 8 + 4], v[i * 8 + 5]),
                         pack_bf16(v[i * 8 + 6], v[i * 8 + 7]));
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)   // 8 x 16B chunks = 32 fp32
            st_shared_v4(rbase + (((uint32_t)i ^ swz) << 4), __float_as_uint(v[i * 4 + 0]),
                         __float_as_uint(v[i * 4 + 1]), __float_as_uint(v[i * 4 + 2]), __float_as_uint(v[i * 4 + 3]));
        }
        // make generic-proxy smem writes visible to the async proxy, then one thread issues the TMA store
        ptx::fence_proxy_async_smem();
        named_bar_sync(bar_id, kGroupThreads);
        if (leader) {
          const void* src = smem_d0 + (sd - ptx::smem_u32(smem_d0));
          if constexpr (REDUCE) ptx::tma_reduce_add_2d(&tmD, src, col0, m0);
          else ptx::tma_store_2d(&tmD, src, col0, m0);
          ptx::tma_commit_group();
        }
        if constexpr (sizeof(OutT) == 2) {
          if (want_stats) {
            // per-column sum / sum-of-squares of the values as stored (rounded), for fused BatchNorm statistics.
            // Warp wq of the group owns column octets {2wq, 2wq+1} of the slab over all 128 rows: lane -> (octet
            // o, 8-row group rg); the row visited at step rr is rotated by rg so the 32 16-byte reads of a warp
            // spread over all 8 swizzle positions (4 wavefronts = the 512-byte minimum).  Rows beyond M are zero
            // because TMA zero-filled A.
            const int wq = gtid >> 5;
            const int o = 2 * wq + (lane & 1), rg = lane >> 1;
            float acc[16];   // [0..8) column sums, [8..16) sums of squares of octet o
#pragma unroll
            for (int k = 0; k < 16; ++k) acc[k] = 0.f;
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
              const uint32_t r8 = (uint32_t)((rr + rg) & 7);
              const uint4 u = ld_shared_v4(sd + (uint32_t)(rg * 8 + r8) * 128u + (((uint32_t)o ^ r8) << 4));
              const uint32_t wds[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float lo = __uint_as_float(wds[k] << 16), hi = __uint_as_float(wds[k] & 0xffff0000u);
                acc[2 * k] += lo; acc[8 + 2 * k] = fmaf(lo, lo, acc[8 + 2 * k]);
                acc[2 * k + 1] += hi; acc[8 + 2 * k + 1] = fmaf(hi, hi, acc[8 + 2 * k + 1]);
              }
            }
            // butterfly that halves the live values at every step: 15 shuffles instead of 64; afterwards lane l
            // holds value index ((l>>4)&1)*8 + ((l>>3)&1)*4 + ((l>>2)&1)*2 + ((l>>1)&1) of octet 2wq + (l&1)
#pragma unroll
            for (int step = 0; step < 4; ++step) {
              const int m = 16 >> step, half = 8 >> step;
              const bool upper = (lane & m) != 0;
#pragma unroll
              for (int i = 0; i < half; ++i) {
                const float send = upper ? acc[i] : acc[i + half];
                const float keep = upper ? acc[i + half] : acc[i];
                acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
              }
            }
            const int vidx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            const int colg = col0 + o * 8 + (vidx & 7);
            if (colg < p.N) {
              if (stats_smem) myacc[(vidx >> 3) * kMaxStatCols + colg] += acc[0];   // one owner lane per column: plain RMW
              else atomicAdd((vidx >> 3) ? (p.col_sumsq + colg) : (p.col_sum + colg), acc[0]);
            }
          }
        }
      }
    }
    if (stats_smem) {
      named_bar_sync(bar_id, kGroupThreads);
      for (int c = gtid; c < p.N; c += kGroupThreads) {
        const float a = myacc[c], b = myacc[kMaxStatCols + c];
        if (a != 0.f || b != 0.f) {
          atomicAdd(p.col_sum + c, a);
          atomicAdd(p.col_sumsq + c, b);
        }
      }
    }
    if (leader) ptx::tma_wait_group<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CTAS == 2) ptx::cluster_sync();   // the leader's MMAs read the peer's shared memory until the last tile
  if (warp == 2) {
    ptx::tc_fence_after();
    if constexpr (CTAS == 2) ptx::tmem_dealloc_2sm(tmem_base, kTmemCols);
    else ptx::tmem_dealloc(tmem_base, kTmemCols);
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
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  if (!fn) throw std::runtime_error("cuTensorMapEncodeTiled not available from the driver");
  return fn;
}

// 2D row-major tensor [rows, cols] with row stride ld (elements); box = [box_rows, box_cols]; 128B swizzle.
static CUtensorMap make_map(const void* ptr, CUtensorMapDataType dt, int elem_bytes, int64_t rows, int64_t cols,
                            int64_t ld, int box_rows, int box_cols) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  // The driver entry point needs a current context; autograd's backward threads may not have touched the
  // runtime yet (CUDA_ERROR_INVALID_CONTEXT otherwise).  cudaFree(0) binds the primary context once per thread.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    CUDA_CHECK(cudaFree(0));
    ctx_bound = true;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) throw std::runtime_error("TMA: base pointer not 16B aligned");
  if ((strides[0] & 15) != 0) throw std::runtime_error("TMA: row stride must be a multiple of 16 bytes");
  CUresult r = get_encode()(&m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return m;
}

template <int BN, bool A_MN, bool B_MN, typename OutT, bool REDUCE, int CTAS = 1>
static void launch(const void* A, const void* B, void* D, int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldd,
                   const float* bias, int relu, float* col_sum, float* col_sumsq, const void* addend, int64_t ldc,
                   int splits, int max_ctas, const int* dbg, cudaStream_t stream) {
  using S = Smem<BN, OutT, CTAS>;
  constexpr int kBRows = BN / CTAS;
  CUtensorMap tmA = A_MN ? make_map(A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, K, M, lda, BLOCK_K, 64)
                         : make_map(A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, K, lda, BLOCK_M, BLOCK_K);
  CUtensorMap tmB = B_MN ? make_map(B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, K, N, ldb, BLOCK_K, 64)
                         : make_map(B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, N, K, ldb, kBRows, BLOCK_K);
  CUtensorMap tmD = sizeof(OutT) == 2 ? make_map(D, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N, ldd, BLOCK_M, 64)
                                      : make_map(D, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, N, ldd, BLOCK_M, 32);
  Params p;
  p.M = M; p.N = N; p.K = K;
  const int nkb = (K + BLOCK_K - 1) / BLOCK_K;
  if (splits < 1) splits = 1;
  if (splits > nkb) splits = nkb;
  p.kb_per_split = (nkb + splits - 1) / splits;
  p.num_splits = (nkb + p.kb_per_split - 1) / p.kb_per_split;
  p.bias = bias; p.relu = relu; p.col_sum = col_sum; p.col_sumsq = col_sumsq;
  p.addend = reinterpret_cast<const __nv_bfloat16*>(addend); p.ldc = ldc;
  if (addend != nullptr && ((reinterpret_cast<uintptr_t>(addend) & 15) != 0 || (ldc % 8) != 0))
    throw std::runtime_error("gemm: addend must be 16B aligned with a row stride multiple of 8");
  if (bias != nullptr && ((reinterpret_cast<uintptr_t>(bias) & 15) != 0 || (N % 4) != 0))
    throw std::runtime_error("gemm: bias must be 16B aligned and N a multiple of 4");
  p.dbg_lbo_a = dbg ? dbg[0] : 0; p.dbg_sbo_a = dbg ? dbg[1] : 0; p.dbg_lbo_b = dbg ? dbg[2] : 0; p.dbg_sbo_b = dbg ? dbg[3] : 0;
  const int num_m = (M + BLOCK_M * CTAS - 1) / (BLOCK_M * CTAS), num_n = (N + BN - 1) / BN;
  p.gm = num_m < 16 ? num_m : 16;
  p.full_groups = num_m / p.gm;
  p.tail_gm = num_m - p.full_groups * p.gm;
  p.d_split = make_fastdiv((uint32_t)p.num_splits);
  p.d_group = make_fastdiv((uint32_t)(p.gm * num_n));
  p.d_gm = make_fastdiv((uint32_t)p.gm);
  p.d_tail = make_fastdiv((uint32_t)(p.tail_gm > 0 ? p.tail_gm : 1));
  const int64_t num_work64 = (int64_t)num_m * num_n * p.num_splits;
  if (num_work64 > (1ll << 30)) throw std::runtime_error("gemm: too many tiles");
  const int num_work = (int)num_work64;
  int clusters = (max_ctas > 0 ? max_ctas : kNumSMs) / CTAS;
  if (clusters > num_work) clusters = num_work;
  if (clusters < 1) clusters = 1;
  auto kern = gemm_kernel<BN, A_MN, B_MN, OutT, REDUCE, CTAS>;
  CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
  if constexpr (CTAS == 1) {
    kern<<<clusters, kThreads, S::kTotal, stream>>>(tmA, tmB, tmD, p);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * CTAS), 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = S::kTotal;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmD, p));
  }
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace gemm
}  // namespace mxb

using namespace mxb;

// mode: 0 = NT (A K-major, B K-major)  -> bf16 out
//       1 = NN (A K-major, B MN-major) -> bf16 out
//       2 = TN (A MN-major, B MN-major) -> fp32 out, TMA reduce-add (caller zero-fills D), split-K
// A, B bf16.  lda/ldb/ldd are row strides (elements) of the *stored* 2D arrays.
MXB_API int MXBGemmBf16(int mode, const void* A, const void* B, void* D, int M, int N, int K, int64_t lda, int64_t ldb,
                        int64_t ldd, const float* bias, int relu, float* col_sum, float* col_sumsq, const void* addend,
                        int64_t ldc, int splits, int block_n, int max_ctas, const int* dbg, void* stream) {
  MXB_TRY
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  MXB_CHECK(M > 0 && N > 0 && K > 0, "empty GEMM");
  // tile width: explicit, or the cheapest by a wave-quantisation estimate (a 256-wide tile costs 2 units, a 128-wide
  // one 1.25: it moves 1.33x the operand bytes per flop through shared memory)
  int bn = block_n;
  if (bn != 64 && bn != 128 && bn != 256) {
    // measured on B200 (profiles/microbench_*): the 256-wide tile wins whenever N > 128, even at worse wave
    // quantisation (it moves 25% fewer operand bytes per flop through shared memory); tiny problems prefer more,
    // narrower tiles so that more SMs take part
    bn = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
    const int64_t nm = (M + 127) / 128;
    const int sp = (mode == 2 && splits > 1) ? splits : 1;
    while (bn > 64 && nm * ((N + bn - 1) / bn) * sp < 100) bn >>= 1;
  }
  // CTA pairs (cta_group::2): 256 x 256 tiles shared by two SMs; worth it when the problem is compute bound (deep K) and
  // has enough 256-row tiles to fill the 74 clusters.  MXB200_GEMM_2CTA=0/1 forces it off/on.
  static const int two_cta_env = [] { const char* e = getenv("MXB200_GEMM_2CTA"); return e ? atoi(e) : -1; }();
  static const int two_cta_wgrad = [] { const char* e = getenv("MXB200_GEMM_2CTA_WGRAD"); return e ? atoi(e) : 1; }();
  bool two_cta = false;
  if (bn == 256 && (mode != 2 || two_cta_wgrad == 1)) {
    const int sp2 = (mode == 2 && splits > 1) ? splits : 1;
    const int64_t tiles2 = (((int64_t)M + 255) / 256) * (((int64_t)N + 255) / 256) * sp2;
    two_cta = two_cta_env == 1 || (two_cta_env != 0 && K >= 256 && tiles2 >= 74);
    if (mode == 2 && M < 256) two_cta = false;
  }
#define MXB_GEMM_DISPATCH(AMN, BMN, OUT, RED, SPL)                                                                   \
  do {                                                                                                                \
    if (two_cta) gemm::launch<256, AMN, BMN, OUT, RED, 2>(A, B, D, M, N, K, lda, ldb, ldd, bias, relu, col_sum, col_sumsq, addend, ldc, SPL, max_ctas, dbg, s); \
    else if (bn == 64) gemm::launch<64, AMN, BMN, OUT, RED>(A, B, D, M, N, K, lda, ldb, ldd, bias, relu, col_sum, col_sumsq, addend, ldc, SPL, max_ctas, dbg, s); \
    else if (bn == 128) gemm::launch<128, AMN, BMN, OUT, RED>(A, B, D, M, N, K, lda, ldb, ldd, bias, relu, col_sum, col_sumsq, addend, ldc, SPL, max_ctas, dbg, s); \
    else gemm::launch<256, AMN, BMN, OUT, RED>(A, B, D, M, N, K, lda, ldb, ldd, bias, relu, col_sum, col_sumsq, addend, ldc, SPL, max_ctas, dbg, s); \
  } while (0)
  if (mode == 0) {
    MXB_GEMM_DISPATCH(false, false, __nv_bfloat16, false, 1);
  } else if (mode == 1) {
    MXB_GEMM_DISPATCH(false, true, __nv_bfloat16, false, 1);
  } else if (mode == 2) {
    MXB_CHECK(bias == nullptr && relu == 0 && col_sum == nullptr && addend == nullptr, "TN mode has a plain reduce-add epilogue");
    MXB_GEMM_DISPATCH(true, true, float, true, splits);
  } else {
    throw std::runtime_error("MXBGemmBf16: bad mode");
  }
  MXB_CATCH
}
