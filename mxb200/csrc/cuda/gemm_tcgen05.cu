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
#include <string.h>

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

// Implicit-GEMM convolution geometry (CONV != 0): one operand is gathered by im2col-mode TMA straight from the NHWC
// activation tensor, so no patch matrix ever exists in memory (the reference materialises it: src/operator/nn/im2col.cuh
// + convolution-inl.h:150-230; cuDNN's implicit GEMM is the library equivalent).
//   CONV == 1 : A = im2col(act)  [pixels x (tap, channel)]  K-major      fwd:   Y  = im2col(X)  . W^T
//                                                                         dgrad: dX = im2col(dY) . W   (flipped taps)
//   CONV == 2 : B = im2col(act)  [(pixels) x (tap, channel)] MN-major    wgrad: dW = dY^T . im2col(X), reduction = pixels
// A "tap" is one filter position; per tap the table gives the pixel offset inside the bounding box and the weight tap it
// pairs with (dgrad flips them; the parity sub-problems of a strided dgrad use a subset).
constexpr int kMaxTaps = 49;
struct ConvGeom {
  int C;              // channels of the gathered tensor (multiple of 64)
  int cpt;            // C / 64: k-blocks per tap
  int ntaps;
  int P, Q;           // base-pixel grid per image (rows, cols): GEMM pixel index = (n * P + p) * Q + q
  int stride;         // traversal stride of the base pixels (the map's elementStrides)
  int base_h, base_w; // lower corner of the bounding box (= -pad for fwd / wgrad)
  int red_per_tap;    // B reduction-coordinate step per weight tap   (fwd:   Cin, weights [Cout][taps*Cin] K-major)
  int out_per_tap;    // B output-column offset per weight tap        (dgrad: Cin, weights read MN-major)
  // CONV == 3 (row-shifted taps): pixels are indexed over PADDED rows of Qp = Q + kw - 1 positions; the kw taps of one
  // filter row read the SAME smem A tile entered off_w[tap] rows (128 B each) lower, so A crosses L2->smem once per
  // filter row instead of once per tap.  Rows with (index % Qp) >= Q are padding positions and are not stored.
  int Qp, kw, Qout;
  int b_resident;     // CONV == 3: every tap's weight tile has its own smem slot, loaded during the CTA's first tile only
  FastDiv d_cpt, d_Q, d_PQ, d_C;
  int8_t off_w[kMaxTaps], off_h[kMaxTaps], wtap[kMaxTaps];
  // strided (parity) dgrad: rows are scattered to dst[((n * dH + p * dstep + ph) * dW + q * dstep + pw) * dC + col]
  __nv_bfloat16* direct;
  int dH, dW, dC, dstep, ph, pw;
};

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
  ConvGeom conv;
};

// CONV == 3: separate rings -- an A stage (136 gathered pixels x 128 B) feeds kw taps, a B stage holds ONE tap's weights.
constexpr int kShiftPixels = 136;
constexpr int kShiftABytes = kShiftPixels * 128;     // 17 KB = 17 x 1024: every stage stays 1024-byte aligned
constexpr int kShiftAStages = 4;
template <int BN>
struct SmemShift {
  static constexpr int kBBytes = BN * BLOCK_K * 2;
  static constexpr int kGroupBufs = 1;
  static constexpr int kStagingBytes = 2 * kGroupBufs * kSlabBytes;
  static constexpr int kStatBytes = 2 * 2 * kMaxStatCols * 4;
  static constexpr int kBarBytes = 512;
  static constexpr int kBudget = 227 * 1024 - 1024 - kBarBytes - kStatBytes - kStagingBytes - kShiftAStages * kShiftABytes;
  static constexpr int kBStagesRaw = kBudget / kBBytes;
  static constexpr int kBStages = kBStagesRaw > 12 ? 12 : kBStagesRaw;
  static constexpr int kTotal = kShiftAStages * kShiftABytes + kBStages * kBBytes + kStagingBytes + kStatBytes + kBarBytes + 1024;
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
template <int BN, bool A_MN, bool B_MN, typename OutT, bool REDUCE, int CTAS, int CONV>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmD, const Params p) {
  using S = Smem<BN, OutT, CTAS>;
  constexpr int kBRows = BN / CTAS;                 // B rows staged by this CTA
  constexpr int kTileM = BLOCK_M * CTAS;            // rows of the output tile owned by the cluster
  const uint32_t cta_rank = CTAS == 2 ? ptx::cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  const int cluster_id = (int)blockIdx.x / CTAS, num_clusters = (int)gridDim.x / CTAS;
  constexpr bool SHIFT = (CONV == 3);               // row-shifted taps: A ring + per-tap B ring (SmemShift)
  using SS = SmemShift<BN>;
  static_assert(!SHIFT || (CTAS == 1 && !A_MN && sizeof(OutT) == 2 && !REDUCE && SS::kBStages >= 4), "shifted-tap convolution: 1 CTA, K-major A, bf16 out");
  constexpr int kStages = SHIFT ? kShiftAStages : S::kStages;       // A ring (both operands share it unless SHIFT)
  constexpr int kBStages = SHIFT ? SS::kBStages : S::kStages;
  constexpr int kABytes = SHIFT ? kShiftABytes : S::kABytes;
  constexpr int kBBytes = SHIFT ? SS::kBBytes : S::kBBytes;
  constexpr int kStagingBytes = SHIFT ? SS::kStagingBytes : S::kStagingBytes;
  constexpr int kTmemCols = 2 * BN;  // double-buffered fp32 accumulators (power of two >= 32)
  constexpr int kSlabCols = S::kSlabCols;
  constexpr int kSlabs = S::kSlabs;
  constexpr int kGroupBufs = SHIFT ? SS::kGroupBufs : S::kGroupBufs;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kABytes;
  uint8_t* smem_d0 = smem_b + kBStages * kBBytes;
  float* colacc = reinterpret_cast<float*>(smem_d0 + kStagingBytes);   // [group][sum|sumsq][kMaxStatCols]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(colacc) + S::kStatBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  uint64_t* bfull_bar = bars + 2 * kStages + 6;              // SHIFT only: the per-tap B ring
  uint64_t* bempty_bar = bfull_bar + kBStages;

  // warp index through a shuffle: the compiler then knows it is warp-uniform, so the role branches are uniform branches and
  // the single-thread roles (TMA producer, MMA issuer) keep their loop state in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
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
    if constexpr (SHIFT) {
      for (int i = 0; i < kBStages; ++i) {
        ptx::mbar_init(&bfull_bar[i], 1);
        ptx::mbar_init(&bempty_bar[i], 1);
      }
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
    // ===================== TMA producer (whole warp walks the loop, the elected lane issues: see the MMA warp) =====================
    {
      const bool issuer = ptx::elect_one();
      int stage = 0;
      uint32_t phase = 0;
      auto load = [&](void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
        if (!issuer) return;
        if constexpr (CTAS == 2) ptx::tma_load_2d_2sm(dst, tm, bar, c0, c1);   // bytes credited to the leader's barrier
        else ptx::tma_load_2d(dst, tm, bar, c0, c1);
      };
      auto load_im2col = [&](void* dst, const CUtensorMap* tm, uint64_t* bar, int c, int wq, int hp, int n, int tap) {
        const uint16_t ow = (uint16_t)p.conv.off_w[tap], oh = (uint16_t)p.conv.off_h[tap];
        if (!issuer) return;
        if constexpr (CTAS == 2) ptx::tma_load_im2col_4d_2sm(dst, tm, bar, c, wq, hp, n, ow, oh);
        else ptx::tma_load_im2col_4d(dst, tm, bar, c, wq, hp, n, ow, oh);
      };
      if constexpr (SHIFT) {
        // one gathered A tile (136 padded-row positions x 64 channels) per (filter row, channel chunk); one B tile per tap
        int bstage = 0;
        uint32_t bphase = 0;
        const int ka_total = num_kb_total / p.conv.kw;
        const bool resident = p.conv.b_resident != 0;       // weight-stationary: B slots are filled once per CTA
        bool first_tile = true;
        for (int w = cluster_id; w < num_work; w += num_clusters) {
          int m_blk, n_blk, kb0, kb1;
          decode(w, m_blk, n_blk, kb0, kb1);
          const int m0 = m_blk * BLOCK_M, n0 = n_blk * BN;
          const uint32_t img = p.conv.d_PQ.div((uint32_t)m0);
          const uint32_t rem = (uint32_t)m0 - img * p.conv.d_PQ.d;
          const uint32_t pr = p.conv.d_Q.div(rem);
          const int a_h = p.conv.base_h + (int)pr, a_w = p.conv.base_w + (int)(rem - pr * p.conv.d_Q.d);
          for (int ka = 0; ka < ka_total; ++ka) {
            const uint32_t fr = p.conv.d_cpt.div((uint32_t)ka);            // filter row
            const int cc = (ka - (int)fr * p.conv.cpt) * BLOCK_K;
            const int t0 = (int)fr * p.conv.kw;
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            if (issuer) ptx::mbar_arrive_expect_tx(&full_bar[stage], kABytes);
            if (issuer) ptx::tma_load_im2col_4d(smem_a + stage * kABytes, &tmA, &full_bar[stage], cc, a_w, a_h, (int)img, (uint16_t)0,
                                    (uint16_t)p.conv.off_h[t0]);
            for (int sx = 0; sx < p.conv.kw; ++sx) {
              const int wt = (int)p.conv.wtap[t0 + sx];
              if (resident) {
                if (!first_tile) continue;
                bstage = ka * p.conv.kw + sx;                 // slot of (filter row, channel chunk, tap in row)
              } else {
                ptx::mbar_wait(&bempty_bar[bstage], bphase ^ 1);
              }
              if (issuer) ptx::mbar_arrive_expect_tx(&bfull_bar[bstage], kBBytes);
              uint8_t* b_dst = smem_b + bstage * kBBytes;
              if constexpr (B_MN) {
#pragma unroll
                for (int c = 0; c < BN / 64; ++c)
                  if (issuer) ptx::tma_load_2d(b_dst + c * (BLOCK_K * 128), &tmB, &bfull_bar[bstage], wt * p.conv.out_per_tap + n0 + c * 64, cc);
              } else {
                if (issuer) ptx::tma_load_2d(b_dst, &tmB, &bfull_bar[bstage], wt * p.conv.red_per_tap + cc, n0);
              }
              if (!resident && ++bstage == kBStages) {
                bstage = 0;
                bphase ^= 1;
              }
            }
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
          first_tile = false;
        }
      } else
      for (int w = cluster_id; w < num_work; w += num_clusters) {
        int m_blk, n_blk, kb0, kb1;
        decode(w, m_blk, n_blk, kb0, kb1);
        const int m0 = m_blk * kTileM + (int)cta_rank * BLOCK_M;     // first A / output row of this CTA
        const int n0 = n_blk * BN + (int)cta_rank * kBRows;          // first B row staged by this CTA
        // CONV == 1: base pixel of the first row of this CTA's A tile
        int a_n = 0, a_h = 0, a_w = 0;
        if constexpr (CONV == 1) {
          const uint32_t img = p.conv.d_PQ.div((uint32_t)m0);
          const uint32_t rem = (uint32_t)m0 - img * p.conv.d_PQ.d;
          const uint32_t pr = p.conv.d_Q.div(rem);
          a_n = (int)img;
          a_h = p.conv.base_h + (int)pr * p.conv.stride;
          a_w = p.conv.base_w + (int)(rem - pr * p.conv.d_Q.d) * p.conv.stride;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (is_leader && issuer) ptx::mbar_arrive_expect_tx(&full_bar[stage], CTAS * S::kStageBytes);
          uint8_t* a_dst = smem_a + stage * kABytes;
          uint8_t* b_dst = smem_b + stage * kBBytes;
          if constexpr (CONV == 1) {
            // k-block kb = (tap, 64-channel chunk)
            const uint32_t tap = p.conv.d_cpt.div((uint32_t)kb);
            const int cc = (kb - (int)tap * p.conv.cpt) * BLOCK_K;
            const int wt = (int)p.conv.wtap[tap];
            load_im2col(a_dst, &tmA, &full_bar[stage], cc, a_w, a_h, a_n, (int)tap);
            if constexpr (B_MN) {
#pragma unroll
              for (int c = 0; c < kBRows / 64; ++c)
                load(b_dst + c * (BLOCK_K * 128), &tmB, &full_bar[stage], wt * p.conv.out_per_tap + n0 + c * 64, cc);
            } else {
              load(b_dst, &tmB, &full_bar[stage], wt * p.conv.red_per_tap + cc, n0);
            }
          } else {
            if constexpr (A_MN) {
              // A stored [K, M]: boxes of [64 k][64 m]; chunk c covers m in [64c, 64c+64)
#pragma unroll
              for (int c = 0; c < BLOCK_M / 64; ++c)
                load(a_dst + c * (BLOCK_K * 128), &tmA, &full_bar[stage], m0 + c * 64, kb * BLOCK_K);
            } else {
              load(a_dst, &tmA, &full_bar[stage], kb * BLOCK_K, m0);
            }
            if constexpr (CONV == 2) {
              // B chunk c = 64 (tap, channel) columns gathered for the 64 pixels of this k-block
              const uint32_t pix = (uint32_t)kb * BLOCK_K;
              const uint32_t img = p.conv.d_PQ.div(pix);
              const uint32_t rem = pix - img * p.conv.d_PQ.d;
              const uint32_t pr = p.conv.d_Q.div(rem);
              const int b_h = p.conv.base_h + (int)pr * p.conv.stride;
              const int b_w = p.conv.base_w + (int)(rem - pr * p.conv.d_Q.d) * p.conv.stride;
#pragma unroll
              for (int c = 0; c < kBRows / 64; ++c) {
                int col = n0 + c * 64;
                if (col >= p.N) col = 0;               // tail tile: any valid data, the store clips these columns
                const uint32_t tap = p.conv.d_C.div((uint32_t)col);
                load_im2col(b_dst + c * (BLOCK_K * 128), &tmB, &full_bar[stage], col - (int)tap * p.conv.C, b_w, b_h, (int)img,
                            (int)tap);
              }
            } else if constexpr (B_MN) {
#pragma unroll
              for (int c = 0; c < kBRows / 64; ++c)
                load(b_dst + c * (BLOCK_K * 128), &tmB, &full_bar[stage], n0 + c * 64, kb * BLOCK_K);
            } else {
              load(b_dst, &tmB, &full_bar[stage], kb * BLOCK_K, n0);
            }
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop (barrier waits, stage / phase bookkeeping) so every operand is provably warp-uniform
    // and lives in uniform registers; only the elected lane issues tcgen05.mma / tcgen05.commit.  (A loop entered by
    // lane 0 alone keeps the descriptors in vector registers: ~30 SASS instructions incl. five R2UR per MMA.)
    if (is_leader) {
      const bool issuer = ptx::elect_one();
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
      int bstage = 0;
      uint32_t bphase = 0;
      int it = 0;
      const uint64_t adesc0 = ptx::make_smem_desc(ptx::smem_u32(smem_a), lbo_a, sbo_a);
      const uint64_t bdesc0 = ptx::make_smem_desc(ptx::smem_u32(smem_b), lbo_b, sbo_b);
      for (int w = cluster_id; w < num_work; w += num_clusters, ++it) {
        int m_blk, n_blk, kb0, kb1;
        decode(w, m_blk, n_blk, kb0, kb1);
        const int buf = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        ptx::mbar_wait(&tmem_empty[buf], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        if constexpr (SHIFT) {
          const int ka_total = num_kb_total / p.conv.kw;
          const bool resident = p.conv.b_resident != 0;
          for (int ka = 0; ka < ka_total; ++ka) {
            const int t0 = (int)p.conv.d_cpt.div((uint32_t)ka) * p.conv.kw;
            ptx::mbar_wait(&full_bar[stage], phase);
            ptx::tc_fence_after();
            const uint32_t a_addr = ptx::smem_u32(smem_a + stage * kABytes);
            for (int sx = 0; sx < p.conv.kw; ++sx) {
              if (resident) {
                bstage = ka * p.conv.kw + sx;
                if (it == 0) {                                  // the slot is written once; later tiles find it in place
                  ptx::mbar_wait(&bfull_bar[bstage], 0);
                  ptx::tc_fence_after();
                }
              } else {
                ptx::mbar_wait(&bfull_bar[bstage], bphase);
                ptx::tc_fence_after();
              }
              // this tap reads the gathered tile entered `shift` positions lower (rows are 128 B apart).  The 128B swizzle
              // is a function of the ABSOLUTE shared-memory address on sm_100 (measured: the descriptor's base-offset field
              // must stay 0; setting it to the row phase corrupts the operand), so a row-shifted start address just works
              const uint32_t shift = (uint32_t)p.conv.off_w[t0 + sx];
              const uint32_t b_addr = ptx::smem_u32(smem_b + bstage * kBBytes);
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                uint64_t adesc = ptx::make_smem_desc(a_addr + shift * 128u + k * kstep_a, lbo_a, sbo_a, p.conv.ph ? shift : 0u);
                uint64_t bdesc = ptx::make_smem_desc(b_addr + k * kstep_b, lbo_b, sbo_b);
                if (issuer) ptx::umma_f16(tmem_d, adesc, bdesc, idesc, (ka > 0 || sx > 0 || k > 0) ? 1u : 0u);
              }
              if (!resident) {
                if (issuer) ptx::umma_commit(&bempty_bar[bstage]);
                if (++bstage == kBStages) {
                  bstage = 0;
                  bphase ^= 1;
                }
              }
            }
            if (issuer) ptx::umma_commit(&empty_bar[stage]);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (issuer) ptx::umma_commit(&tmem_full[buf]);
          continue;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          // descriptors = stage-0 descriptor + byte offset >> 4 in the 14-bit address field (shared memory is < 256 KB, the
          // field cannot carry out): two 64-bit adds per k-block and one per MMA instead of rebuilding the bit fields
          const uint64_t adesc_s = adesc0 + (uint64_t)((uint32_t)stage * (uint32_t)(kABytes >> 4));
          const uint64_t bdesc_s = bdesc0 + (uint64_t)((uint32_t)stage * (uint32_t)(kBBytes >> 4));
          if (issuer) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint64_t adesc = adesc_s + (uint64_t)((uint32_t)k * (kstep_a >> 4));
              const uint64_t bdesc = bdesc_s + (uint64_t)((uint32_t)k * (kstep_b >> 4));
              if constexpr (CTAS == 2) ptx::umma_f16_2sm(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              else ptx::umma_f16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            // frees the smem slot (in both CTAs of a pair) when these MMAs retire
            if constexpr (CTAS == 2) ptx::umma_commit_2sm(&empty_bar[stage], 3);
            else ptx::umma_commit(&empty_bar[stage]);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        // accumulator ready for the epilogue (of both CTAs)
        if (issuer) {
          if constexpr (CTAS == 2) ptx::umma_commit_2sm(&tmem_full[buf], 3);
          else ptx::umma_commit(&tmem_full[buf]);
        }
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
      // SHIFT: rows are positions of PADDED image rows; only the first Qout of every Qp positions are output pixels
      bool row_valid = grow < p.M;
      int64_t out_pix = grow;
      if constexpr (SHIFT) {
        const uint32_t prow = p.conv.d_Q.div((uint32_t)grow);            // d_Q divides by Qp in this mode
        const uint32_t qp = (uint32_t)grow - prow * p.conv.d_Q.d;
        row_valid = row_valid && (int)qp < p.conv.Qout;
        out_pix = (int64_t)prow * p.conv.Qout + qp;
      }
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
            const __nv_bfloat16* src = p.addend + out_pix * p.ldc + col0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (row_valid && col0 + i * 8 < p.N) cadd[i] = *reinterpret_cast<const uint4*>(src + i * 8);
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
        bool direct = false;
        if constexpr (CONV == 1 && sizeof(OutT) == 2) direct = p.conv.direct != nullptr;
        if constexpr (SHIFT) direct = true;
        if (direct) {
          if constexpr ((CONV == 1 || SHIFT) && sizeof(OutT) == 2) {
            // this thread's pixel goes straight to global memory (8 x 16 B = one full 128-byte line per thread and slab):
            // strided-dgrad parity planes scatter to a strided position, shifted-tap tiles drop the padding positions
            if (row_valid) {
              __nv_bfloat16* dst;
              if constexpr (SHIFT) {
                dst = p.conv.direct + out_pix * p.conv.dC + col0;
              } else {
                const uint32_t img = p.conv.d_PQ.div((uint32_t)grow);
                const uint32_t rem = (uint32_t)grow - img * p.conv.d_PQ.d;
                const uint32_t pr = p.conv.d_Q.div(rem);
                const uint32_t qc = rem - pr * p.conv.d_Q.d;
                dst = p.conv.direct +
                      (((int64_t)img * p.conv.dH + (int64_t)pr * p.conv.dstep + p.conv.ph) * p.conv.dW + (int64_t)qc * p.conv.dstep +
                       p.conv.pw) * p.conv.dC + col0;
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (col0 + i * 8 < p.N)
                  *reinterpret_cast<uint4*>(dst + i * 8) =
                      make_uint4(pack_bf16(v[i * 8 + 0], v[i * 8 + 1]), pack_bf16(v[i * 8 + 2], v[i * 8 + 3]),
                                 pack_bf16(v[i * 8 + 4], v[i * 8 + 5]), pack_bf16(v[i * 8 + 6], v[i * 8 + 7]));
              }
            }
          }
          if (!(SHIFT && want_stats)) continue;
          // statistics of a shifted-tap tile: stage the tile (padding positions as zeros) for the column sums below
          if (!row_valid) {
#pragma unroll
            for (int j = 0; j < kSlabCols; ++j) v[j] = 0.f;
          }
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
        if (leader && !direct) {
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

// The tensor-map encoders are DRIVER entry points and need a current context on the calling thread; autograd's backward
// threads may not have one yet (CUDA_ERROR_INVALID_CONTEXT otherwise).  Bind the device's primary context through the
// driver API: unlike the usual cudaFree(0) idiom this does not synchronise the legacy stream, which would invalidate a
// CUDA-graph capture in progress on another stream (parallel/graph.py).
static void bind_context() {
  static thread_local bool bound = false;
  if (bound) return;
  typedef CUresult (*GetCur)(CUcontext*);
  typedef CUresult (*SetCur)(CUcontext);
  typedef CUresult (*Retain)(CUcontext*, CUdevice);
  static GetCur get_cur = nullptr;
  static SetCur set_cur = nullptr;
  static Retain retain = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuCtxGetCurrent", &f, cudaEnableDefault, &q) == cudaSuccess) get_cur = reinterpret_cast<GetCur>(f);
    if (cudaGetDriverEntryPoint("cuCtxSetCurrent", &f, cudaEnableDefault, &q) == cudaSuccess) set_cur = reinterpret_cast<SetCur>(f);
    if (cudaGetDriverEntryPoint("cuDevicePrimaryCtxRetain", &f, cudaEnableDefault, &q) == cudaSuccess) retain = reinterpret_cast<Retain>(f);
  });
  CUcontext cur = nullptr;
  if (get_cur && get_cur(&cur) == CUDA_SUCCESS && cur != nullptr) {
    bound = true;
    return;
  }
  int dev = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  if (retain && set_cur && retain(&cur, (CUdevice)dev) == CUDA_SUCCESS && set_cur(cur) == CUDA_SUCCESS) {
    bound = true;
    return;
  }
  CUDA_CHECK(cudaFree(0));   // last resort
  bound = true;
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
  bind_context();
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) throw std::runtime_error("TMA: base pointer not 16B aligned");
  if ((strides[0] & 15) != 0) throw std::runtime_error("TMA: row stride must be a multiple of 16 bytes");
  CUresult r = get_encode()(&m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return m;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

static EncodeIm2colFn get_encode_im2col() {
  static EncodeIm2colFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeIm2colFn>(f);
  });
  if (!fn) throw std::runtime_error("cuTensorMapEncodeIm2col not available from the driver");
  return fn;
}

// Host description of the gathered activation tensor of an implicit-GEMM convolution.
struct ConvDesc {
  const void* act;              // NHWC bf16
  int n, h, w, c;
  int lower_h, lower_w;         // bounding box of the base pixels: [lower, dim + upper)
  int upper_h, upper_w;
  int64_t b_rows, b_cols;       // dgrad: the weight matrix as stored ([Cout][taps * Cin]) -- differs from the GEMM's K x N
  ConvGeom geom;                // device geometry (fast divisors are filled in by launch)
};

// im2col-mode tensor map over an NHWC bf16 tensor: ``pixels`` base pixels x 64 channels per load, 128B swizzle.
static CUtensorMap make_im2col_map(const ConvDesc& cd, int pixels) {
  CUtensorMap m;
  bind_context();
  cuuint64_t dims[4] = {(cuuint64_t)cd.c, (cuuint64_t)cd.w, (cuuint64_t)cd.h, (cuuint64_t)cd.n};
  cuuint64_t strides[3] = {(cuuint64_t)cd.c * 2, (cuuint64_t)cd.w * cd.c * 2, (cuuint64_t)cd.h * cd.w * cd.c * 2};
  int lower[2] = {cd.lower_w, cd.lower_h};
  int upper[2] = {cd.upper_w, cd.upper_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)cd.geom.stride, (cuuint32_t)cd.geom.stride, 1};
  if ((reinterpret_cast<uintptr_t>(cd.act) & 15) != 0) throw std::runtime_error("im2col TMA: base pointer not 16B aligned");
  if (cd.c % 64 != 0) throw std::runtime_error("im2col TMA: channels must be a multiple of 64");
  CUresult r = get_encode_im2col()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(cd.act), dims, strides, lower,
                                   upper, 64, (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeIm2col failed with code " + std::to_string((int)r));
  // drivers <= 13.1 set a descriptor bit that is wrong for tensors smaller than 128 KB (same work-around as CUTLASS,
  // cute/atom/copy_traits_sm90_im2col.hpp)
  static const int drv = [] { int v = 0; cudaDriverGetVersion(&v); return v; }();
  if (drv <= 13010 && (int64_t)cd.n * cd.h * cd.w * cd.c * 2 < 131072) reinterpret_cast<uint64_t*>(&m)[1] &= ~(1ull << 21);
  return m;
}

template <int BN, bool A_MN, bool B_MN, typename OutT, bool REDUCE, int CTAS = 1, int CONV = 0>
static void launch(const void* A, const void* B, void* D, int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldd,
                   const float* bias, int relu, float* col_sum, float* col_sumsq, const void* addend, int64_t ldc,
                   int splits, int max_ctas, const int* dbg, cudaStream_t stream, const ConvDesc* cd = nullptr) {
  using S = Smem<BN, OutT, CTAS>;
  constexpr int kBRows = BN / CTAS;
  static_assert(CONV == 0 || ((CONV == 1 || CONV == 3) && !A_MN) || (CONV == 2 && A_MN && B_MN), "conv operand layouts");
  if (CONV != 0 && cd == nullptr) throw std::runtime_error("conv launch without geometry");
  CUtensorMap tmA = CONV == 3 ? make_im2col_map(*cd, kShiftPixels) : CONV == 1 ? make_im2col_map(*cd, BLOCK_M) : A_MN ? make_map(A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, K, M, lda, BLOCK_K, 64)
                         : make_map(A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, K, lda, BLOCK_M, BLOCK_K);
  const int64_t b_red = ((CONV == 1 || CONV == 3) && cd->b_rows) ? cd->b_rows : K, b_out = ((CONV == 1 || CONV == 3) && cd->b_cols) ? cd->b_cols : N;
  CUtensorMap tmB = CONV == 2 ? make_im2col_map(*cd, BLOCK_K) : B_MN ? make_map(B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, b_red, b_out, ldb, BLOCK_K, 64)
                         : make_map(B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, N, K, ldb, kBRows, BLOCK_K);
  CUtensorMap tmD = sizeof(OutT) == 2 ? make_map(D, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N, ldd, BLOCK_M, 64)
                                      : make_map(D, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, N, ldd, BLOCK_M, 32);
  Params p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  if constexpr (CONV != 0) {
    p.conv = cd->geom;
    p.conv.cpt = p.conv.C / BLOCK_K;
    p.conv.d_cpt = make_fastdiv((uint32_t)p.conv.cpt);
    const int row_len = CONV == 3 ? p.conv.Qp : p.conv.Q;
    p.conv.d_Q = make_fastdiv((uint32_t)row_len);
    p.conv.d_PQ = make_fastdiv((uint32_t)(p.conv.P * row_len));
    if constexpr (CONV == 3) {
      p.conv.direct = reinterpret_cast<__nv_bfloat16*>(D);
      p.conv.dC = (int)ldd;
    }
    p.conv.d_C = make_fastdiv((uint32_t)p.conv.C);
  }
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
  // raster group height: the concurrently running tiles (one per CTA / CTA pair) should cover a near-square patch of the
  // output so that each wave streams the fewest A + B panels (MXB200_GEMM_GM overrides: experiments)
  int gm_want = 16;
  {
    static const int env_gm = [] { const char* e = getenv("MXB200_GEMM_GM"); return e ? atoi(e) : 0; }();
    if (env_gm > 0) gm_want = env_gm;
  }
  p.gm = num_m < gm_want ? num_m : gm_want;
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
  auto kern = gemm_kernel<BN, A_MN, B_MN, OutT, REDUCE, CTAS, CONV>;
  constexpr int kSmemTotal = CONV == 3 ? SmemShift<BN>::kTotal : S::kTotal;
  CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
  if constexpr (CTAS == 1) {
    kern<<<clusters, kThreads, kSmemTotal, stream>>>(tmA, tmB, tmD, p);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * CTAS), 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = kSmemTotal;
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
  // 256 x 128 CTA-pair tiles (experiment switch MXB200_GEMM_2CTA_128=1, read per call): with N = 128 a single CTA's MMA is
  // bound by its shared-memory operand reads (A 4 KB + B 4 KB per 64 tensor cycles); in a pair each CTA stages half of B
  bool pair128 = false;
  if (bn == 128 && mode != 2 && M >= 256 && K >= 128) {
    const char* e = getenv("MXB200_GEMM_2CTA_128");
    pair128 = e != nullptr && atoi(e) == 1;
  }
#define MXB_GEMM_DISPATCH(AMN, BMN, OUT, RED, SPL)                                                                   \
  do {                                                                                                                \
    if (pair128) gemm::launch<128, AMN, BMN, OUT, RED, 2>(A, B, D, M, N, K, lda, ldb, ldd, bias, relu, col_sum, col_sumsq, addend, ldc, SPL, max_ctas, dbg, s); \
    else if (two_cta) gemm::launch<256, AMN, BMN, OUT, RED, 2>(A, B, D, M, N, K, lda, ldb, ldd, bias, relu, col_sum, col_sumsq, addend, ldc, SPL, max_ctas, dbg, s); \
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


// ------------------------------------------------------------------ implicit-GEMM convolution (NHWC bf16, weights KRSC)
// Reference operators: Convolution forward / backward (src/operator/nn/convolution.cu:40-128 -> cudnn_convolution-inl.h or
// convolution-inl.h im2col + GEMM).  Here the im2col gather is done by the TMA unit (im2col-mode tensor maps) inside the
// tcgen05 GEMM's producer warp: no patch matrix, padding = out-of-bounds zero fill, stride = traversal stride.
namespace {

int conv_out(int in, int k, int stride, int pad) { return (in + 2 * pad - k) / stride + 1; }

int pick_bn(int64_t M, int N, int sp) {
  int bn = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  const int64_t nm = (M + 127) / 128;
  while (bn > 64 && nm * ((N + bn - 1) / bn) * sp < 100) bn >>= 1;
  return bn;
}

bool pick_two_cta(int bn, int64_t M, int N, int K, int sp) {
  static const int env = [] { const char* e = getenv("MXB200_GEMM_2CTA"); return e ? atoi(e) : -1; }();
  if (bn != 256) return false;
  const int64_t tiles2 = ((M + 255) / 256) * (((int64_t)N + 255) / 256) * sp;
  return env == 1 || (env != 0 && K >= 256 && tiles2 >= 74);
}

// Row-shifted taps (CONV == 3): the kw taps of a filter row share ONE gathered A tile, which cuts the A-operand L2 -> smem
// traffic of narrow layers (N <= 128) three-fold.  Measured on B200 (profiles/microbench_conv_r2_call53.json) it is
// nevertheless SLOWER than the per-tap gather (C=64: 360 vs 280 us, C=128: 190 vs 151 us): with N = 64 / 128 the SS-mode
// MMA is bound by its own shared-memory operand reads, and an operand entered at a row offset straddles two swizzle
// atoms per 8-row core matrix.  Kept as an opt-in (MXB200_CONV_SHIFT=1, exercised by tests/test_gpu_kernels.py); the
// profitable structure for these layers is weight-stationary with the pixels on the N side.
// Returns 0 (off), 1 (B ring) or 2 (weight-stationary when every tap's weight tile fits in the B slots: n_out <= 64 and
// taps * C / 64 <= 12, i.e. the 3x3 C = 64 layers; B ring otherwise).
int use_shift(int n_out, int kw, int stride) {
  const char* e = getenv("MXB200_CONV_SHIFT");
  if (e == nullptr || atoi(e) < 1 || stride != 1 || kw < 2 || kw > 9) return 0;
  return n_out <= 128 ? atoi(e) : 0;
}
bool shift_resident(int mode, int n_out, int ntaps, int c) {
  return mode >= 2 && n_out <= 64 && ntaps * (c / 64) <= gemm::SmemShift<64>::kBStages;
}

// debug switch: MXB200_SHIFT_BASEOFF=1 puts the row phase into the descriptor's base-offset field (wrong on sm_100: kept to
// document the experiment, tests/test_gpu_kernels.py passes only with the default 0)
int shift_no_base_offset() {
  const char* e = getenv("MXB200_SHIFT_BASEOFF");
  return (e && atoi(e) == 1) ? 1 : 0;
}

#define MXB_CONV_SHIFT_DISPATCH(BMN, MM, NN, KK, LDA, LDB, LDD)                                                        \
  do {                                                                                                                \
    if (NN <= 64) gemm::launch<64, false, BMN, __nv_bfloat16, false, 1, 3>(A, B, D, MM, NN, KK, LDA, LDB, LDD, bias, relu, col_sum, col_sumsq, addend, ldc, 1, 0, nullptr, s, &cd); \
    else gemm::launch<128, false, BMN, __nv_bfloat16, false, 1, 3>(A, B, D, MM, NN, KK, LDA, LDB, LDD, bias, relu, col_sum, col_sumsq, addend, ldc, 1, 0, nullptr, s, &cd); \
  } while (0)

void check_conv_args(int c, int o, int kh, int kw, int stride) {
  MXB_CHECK(c % 64 == 0, "conv: gathered channel count must be a multiple of 64");
  MXB_CHECK(o % 8 == 0, "conv: channel counts must be multiples of 8");
  MXB_CHECK(kh * kw <= gemm::kMaxTaps && kh >= 1 && kw >= 1, "conv: filter too large");
  MXB_CHECK(stride >= 1 && stride <= 8, "conv: bad stride");
}

}  // namespace

#define MXB_CONV_DISPATCH(CONVK, AMN, BMN, OUT, RED, SPL, MM, NN, KK, LDA, LDB, LDD)                                  \
  do {                                                                                                                \
    if (two_cta) gemm::launch<256, AMN, BMN, OUT, RED, 2, CONVK>(A, B, D, MM, NN, KK, LDA, LDB, LDD, bias, relu, col_sum, col_sumsq, addend, ldc, SPL, 0, nullptr, s, &cd); \
    else if (bn == 64) gemm::launch<64, AMN, BMN, OUT, RED, 1, CONVK>(A, B, D, MM, NN, KK, LDA, LDB, LDD, bias, relu, col_sum, col_sumsq, addend, ldc, SPL, 0, nullptr, s, &cd); \
    else if (bn == 128) gemm::launch<128, AMN, BMN, OUT, RED, 1, CONVK>(A, B, D, MM, NN, KK, LDA, LDB, LDD, bias, relu, col_sum, col_sumsq, addend, ldc, SPL, 0, nullptr, s, &cd); \
    else gemm::launch<256, AMN, BMN, OUT, RED, 1, CONVK>(A, B, D, MM, NN, KK, LDA, LDB, LDD, bias, relu, col_sum, col_sumsq, addend, ldc, SPL, 0, nullptr, s, &cd); \
  } while (0)

// y[n,P,Q,o] = conv(x[n,h,w,c], w[o,kh,kw,c]) (+bias)(+addend)(relu); col_sum / col_sumsq: fused BatchNorm statistics of y.
MXB_API int MXBConvFwdBf16(const void* x, const void* w, void* y, int n, int h, int wd, int c, int o, int kh, int kw, int stride,
                           int pad, const float* bias, int relu, float* col_sum, float* col_sumsq, const void* addend_in,
                           int block_n, void* stream) {
  MXB_TRY
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  check_conv_args(c, o, kh, kw, stride);
  const int P = conv_out(h, kh, stride, pad), Q = conv_out(wd, kw, stride, pad);
  MXB_CHECK(P > 0 && Q > 0 && n > 0, "conv: empty output");
  const int64_t M64 = (int64_t)n * P * Q;
  MXB_CHECK(M64 < (1ll << 31), "conv: too many output pixels");
  const int M = (int)M64, N = o, K = kh * kw * c;
  gemm::ConvDesc cd;
  memset(&cd, 0, sizeof(cd));
  cd.act = x; cd.n = n; cd.h = h; cd.w = wd; cd.c = c;
  cd.lower_h = -pad; cd.lower_w = -pad;
  cd.upper_h = pad - (kh - 1); cd.upper_w = pad - (kw - 1);
  // the bounding box must hold exactly P x Q base pixels
  cd.upper_h -= (h + cd.upper_h - cd.lower_h - 1) - (P - 1) * stride;
  cd.upper_w -= (wd + cd.upper_w - cd.lower_w - 1) - (Q - 1) * stride;
  gemm::ConvGeom& g = cd.geom;
  g.C = c; g.ntaps = kh * kw; g.P = P; g.Q = Q; g.stride = stride; g.base_h = -pad; g.base_w = -pad;
  g.red_per_tap = c; g.out_per_tap = 0;
  for (int r = 0; r < kh; ++r)
    for (int q = 0; q < kw; ++q) {
      const int t = r * kw + q;
      g.off_h[t] = (int8_t)r; g.off_w[t] = (int8_t)q; g.wtap[t] = (int8_t)t;
    }
  const void *A = x, *B = w;
  void* D = y;
  const void* addend = addend_in;          // optional residual [n,P,Q,o] added in the epilogue (inference)
  const int64_t ldc = o;
  if (use_shift(N, kw, stride) && block_n == 0) {
    const int Qp = Q + kw - 1;
    MXB_CHECK((int64_t)n * P * Qp < (1ll << 31), "conv: too many padded pixels");
    cd.upper_w = Qp - wd + cd.lower_w;                 // the bounding box holds Qp positions per row
    g.Qp = Qp; g.kw = kw; g.Qout = Q;
    g.ph = shift_no_base_offset();
    g.b_resident = shift_resident(use_shift(N, kw, stride), N, kh * kw, c) ? 1 : 0;
    for (int t = 0; t < kh * kw; ++t) g.off_w[t] = (int8_t)(t % kw);      // row shift of the tap inside the gathered tile
    const int Ms = n * P * Qp;
    MXB_CONV_SHIFT_DISPATCH(false, Ms, N, K, (int64_t)c, (int64_t)K, (int64_t)o);
    return 0;
  }
  int bn = (block_n == 64 || block_n == 128 || block_n == 256) ? block_n : pick_bn(M, N, 1);
  const bool two_cta = pick_two_cta(bn, M, N, K, 1);
  MXB_CONV_DISPATCH(1, false, false, __nv_bfloat16, false, 1, M, N, K, (int64_t)c, (int64_t)K, (int64_t)o);
  MXB_CATCH
}

// dx[n,h,w,c] = sum_{r,s,k} dy[n,(h+pad-r)/stride,(w+pad-s)/stride,k] * w[k,r,s,c]  (+ addend[n,h,w,c], stride 1 only)
MXB_API int MXBConvDgradBf16(const void* dy, const void* w, void* dx, int n, int h, int wd, int c, int o, int kh, int kw,
                             int stride, int pad, const void* addend_in, void* stream) {
  MXB_TRY
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  check_conv_args(o, c, kh, kw, stride);      // the gathered tensor is dy: o channels
  const int P = conv_out(h, kh, stride, pad), Q = conv_out(wd, kw, stride, pad);
  MXB_CHECK(P > 0 && Q > 0 && n > 0, "conv: empty output");
  MXB_CHECK((int64_t)n * h * wd < (1ll << 31), "conv: too many pixels");
  const float* bias = nullptr;
  const int relu = 0;
  float *col_sum = nullptr, *col_sumsq = nullptr;
  const void *A = dy, *B = w;
  const int64_t ldb = (int64_t)kh * kw * c;
  if (stride == 1) {
    const int M = n * h * wd, N = c, K = kh * kw * o;
    gemm::ConvDesc cd;
    memset(&cd, 0, sizeof(cd));
    cd.act = dy; cd.n = n; cd.h = P; cd.w = Q; cd.c = o;
    cd.lower_h = pad - (kh - 1); cd.lower_w = pad - (kw - 1);
    cd.upper_h = h - P + cd.lower_h; cd.upper_w = wd - Q + cd.lower_w;
    cd.b_rows = o; cd.b_cols = ldb;
    gemm::ConvGeom& g = cd.geom;
    g.C = o; g.ntaps = kh * kw; g.P = h; g.Q = wd; g.stride = 1; g.base_h = cd.lower_h; g.base_w = cd.lower_w;
    g.red_per_tap = 0; g.out_per_tap = c;
    for (int r = 0; r < kh; ++r)
      for (int q = 0; q < kw; ++q) {
        const int t = r * kw + q;
        g.off_h[t] = (int8_t)(kh - 1 - r); g.off_w[t] = (int8_t)(kw - 1 - q); g.wtap[t] = (int8_t)t;
      }
    void* D = dx;
    const void* addend = addend_in;
    const int64_t ldc = c;
    if (use_shift(N, kw, 1)) {
      const int Qp = wd + kw - 1;
      MXB_CHECK((int64_t)n * h * Qp < (1ll << 31), "conv: too many padded pixels");
      cd.upper_w = Qp - Q + cd.lower_w;
      g.Qp = Qp; g.kw = kw; g.Qout = wd;
      g.ph = shift_no_base_offset();
      g.b_resident = shift_resident(use_shift(N, kw, 1), N, kh * kw, o) ? 1 : 0;
      const int Ms = n * h * Qp;
      MXB_CONV_SHIFT_DISPATCH(true, Ms, N, K, (int64_t)o, ldb, (int64_t)c);
      return 0;
    }
    const int bn = pick_bn(M, N, 1);
    const bool two_cta = pick_two_cta(bn, M, N, K, 1);
    MXB_CONV_DISPATCH(1, false, true, __nv_bfloat16, false, 1, M, N, K, (int64_t)o, ldb, (int64_t)c);
  } else {
    MXB_CHECK(addend_in == nullptr, "conv dgrad: the addend epilogue needs stride 1");
    // one sub-convolution per output parity plane: only the taps whose input position is integral contribute
    bool need_zero = false;
    for (int ph = 0; ph < stride && !need_zero; ++ph) {
      int cnt = 0;
      for (int r = 0; r < kh; ++r) cnt += ((ph + pad - r) % stride == 0);
      need_zero = cnt == 0;
    }
    for (int pw = 0; pw < stride && !need_zero; ++pw) {
      int cnt = 0;
      for (int q = 0; q < kw; ++q) cnt += ((pw + pad - q) % stride == 0);
      need_zero = cnt == 0;
    }
    if (need_zero) CUDA_CHECK(cudaMemsetAsync(dx, 0, (size_t)n * h * wd * c * 2, s));
    for (int ph = 0; ph < stride; ++ph)
      for (int pw = 0; pw < stride; ++pw) {
        const int H2 = (h - ph + stride - 1) / stride, W2 = (wd - pw + stride - 1) / stride;
        if (H2 <= 0 || W2 <= 0) continue;
        int dh[gemm::kMaxTaps], dw[gemm::kMaxTaps], wt[gemm::kMaxTaps], nt = 0, min_h = 1 << 20, min_w = 1 << 20;
        for (int r = 0; r < kh; ++r) {
          if ((ph + pad - r) % stride != 0) continue;
          for (int q = 0; q < kw; ++q) {
            if ((pw + pad - q) % stride != 0) continue;
            dh[nt] = (ph + pad - r) / stride; dw[nt] = (pw + pad - q) / stride; wt[nt] = r * kw + q;
            min_h = dh[nt] < min_h ? dh[nt] : min_h; min_w = dw[nt] < min_w ? dw[nt] : min_w;
            ++nt;
          }
        }
        if (nt == 0) continue;
        const int M = n * H2 * W2, N = c, K = nt * o;
        gemm::ConvDesc cd;
        memset(&cd, 0, sizeof(cd));
        cd.act = dy; cd.n = n; cd.h = P; cd.w = Q; cd.c = o;
        cd.lower_h = min_h; cd.lower_w = min_w;
        cd.upper_h = H2 - P + min_h; cd.upper_w = W2 - Q + min_w;
        cd.b_rows = o; cd.b_cols = ldb;
        gemm::ConvGeom& g = cd.geom;
        g.C = o; g.ntaps = nt; g.P = H2; g.Q = W2; g.stride = 1; g.base_h = min_h; g.base_w = min_w;
        g.red_per_tap = 0; g.out_per_tap = c;
        for (int t = 0; t < nt; ++t) {
          g.off_h[t] = (int8_t)(dh[t] - min_h); g.off_w[t] = (int8_t)(dw[t] - min_w); g.wtap[t] = (int8_t)wt[t];
        }
        g.direct = reinterpret_cast<__nv_bfloat16*>(dx);
        g.dH = h; g.dW = wd; g.dC = c; g.dstep = stride; g.ph = ph; g.pw = pw;
        void* D = dx;                      // only used to build a (never used) store map
        const void* addend = nullptr;
        const int64_t ldc = 0;
        const int bn = pick_bn(M, N, 1);
        const bool two_cta = pick_two_cta(bn, M, N, K, 1);
        MXB_CONV_DISPATCH(1, false, true, __nv_bfloat16, false, 1, M, N, K, (int64_t)o, ldb, (int64_t)c);
      }
  }
  MXB_CATCH
}

// dw[o, kh*kw*c] (fp32, ACCUMULATED: caller zero-fills or passes a gradient-arena view) += dy^T . im2col(x)
MXB_API int MXBConvWgradBf16(const void* dy, const void* x, void* dw, int n, int h, int wd, int c, int o, int kh, int kw,
                             int stride, int pad, int splits, int block_n, void* stream) {
  MXB_TRY
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  check_conv_args(c, o, kh, kw, stride);
  const int P = conv_out(h, kh, stride, pad), Q = conv_out(wd, kw, stride, pad);
  MXB_CHECK(P > 0 && Q > 0 && n > 0, "conv: empty output");
  const int64_t K64 = (int64_t)n * P * Q;
  MXB_CHECK(K64 < (1ll << 31), "conv: too many output pixels");
  const int M = o, N = kh * kw * c, K = (int)K64;
  gemm::ConvDesc cd;
  memset(&cd, 0, sizeof(cd));
  cd.act = x; cd.n = n; cd.h = h; cd.w = wd; cd.c = c;
  cd.lower_h = -pad; cd.lower_w = -pad;
  cd.upper_h = pad - (kh - 1); cd.upper_w = pad - (kw - 1);
  cd.upper_h -= (h + cd.upper_h - cd.lower_h - 1) - (P - 1) * stride;
  cd.upper_w -= (wd + cd.upper_w - cd.lower_w - 1) - (Q - 1) * stride;
  gemm::ConvGeom& g = cd.geom;
  g.C = c; g.ntaps = kh * kw; g.P = P; g.Q = Q; g.stride = stride; g.base_h = -pad; g.base_w = -pad;
  for (int r = 0; r < kh; ++r)
    for (int q = 0; q < kw; ++q) {
      const int t = r * kw + q;
      g.off_h[t] = (int8_t)r; g.off_w[t] = (int8_t)q; g.wtap[t] = (int8_t)t;
    }
  const void *A = dy, *B = x;
  void* D = dw;
  const float* bias = nullptr;
  const int relu = 0;
  float *col_sum = nullptr, *col_sumsq = nullptr;
  const void* addend = nullptr;
  const int64_t ldc = 0;
  int bn = (block_n == 64 || block_n == 128 || block_n == 256) ? block_n : (N <= 64 ? 64 : (N < 256 ? 128 : 256));
  if (splits < 1) {
    const int64_t tiles = (((int64_t)M + 127) / 128) * ((N + bn - 1) / bn);
    const int64_t kb = (K64 + 63) / 64;
    int64_t want = (2 * 148) / (tiles > 0 ? tiles : 1);
    if (want < 1) want = 1;
    const int64_t cap = kb >= 8 ? kb / 4 : 1;
    splits = (int)(want < cap ? want : cap);
    if (splits < 1) splits = 1;
  }
  bool two_cta = pick_two_cta(bn, M, N, K, splits) && M >= 256;
  MXB_CONV_DISPATCH(2, true, true, float, true, splits, M, N, K, (int64_t)o, (int64_t)c, (int64_t)N);
  MXB_CATCH
}
