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

#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace mxb {
namespace gemm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int kThreads = 384;      // 4 control warps + 8 epilogue warps
constexpr int kEpiThreads = 256;   // two warps per TMEM lane quarter, each takes half of the tile's columns
constexpr int kMaxStatCols = 2048; // per-CTA column-statistics accumulators kept in shared memory

struct Params {
  int M, N, K;          // logical problem: D[M,N] = A[M,K] * B[N,K]^T   (K = reduction)
  int num_splits;       // split-K factor (>1 only with REDUCE epilogue)
  int kb_per_split;     // k-blocks per split
  const float* bias;    // [N] or null
  int relu;
  float* col_sum;       // [N] fp32 accumulators (atomicAdd) or null : sum over rows of stored output
  float* col_sumsq;     // [N]
  const __nv_bfloat16* addend;  // optional C[M,N] (row stride ldc): D = A*B + C  (residual-gradient fusion)
  int64_t ldc;
  // debug overrides for descriptor fields (0 = use computed); lets one GPU run probe encodings
  int dbg_lbo_a, dbg_sbo_a, dbg_lbo_b, dbg_sbo_b;
};

template <int BN, typename OutT>
struct Smem {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BN * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kDBytes = BLOCK_M * BN * (int)sizeof(OutT);
  static constexpr int kDBufs = sizeof(OutT) == 2 ? 2 : 1;   // double-buffered staging for the streaming bf16 case
  static constexpr int kStatBytes = 2 * kMaxStatCols * 4;
  static constexpr int kBudget = 227 * 1024 - 1024 /*align slack*/ - 256 /*barriers*/ - kStatBytes;
  static constexpr int kStagesRaw = (kBudget - kDBufs * kDBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTotal = kStages * kStageBytes + kDBufs * kDBytes + kStatBytes + 256 + 1024;
};

template <int BN, bool A_MN, bool B_MN, typename OutT, bool REDUCE>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmD, const Params p) {
  using S = Smem<BN, OutT>;
  constexpr int kStages = S::kStages;
  constexpr int kTmemCols = 2 * BN;  // double-buffered fp32 accumulators (power of two >= 32)
  constexpr int kOutBoxCols = 128 / (int)sizeof(OutT);  // columns per 128B-swizzled output box
  constexpr int kOutBoxes = BN / kOutBoxCols;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * S::kABytes;
  uint8_t* smem_d0 = smem + kStages * S::kStageBytes;
  float* colacc = reinterpret_cast<float*>(smem_d0 + S::kDBufs * S::kDBytes);   // [2][kMaxStatCols]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(colacc) + S::kStatBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (p.M + BLOCK_M - 1) / BLOCK_M;
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
      ptx::mbar_init(&tmem_empty[i], kEpiThreads / 32);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_base_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  auto decode = [&](int w, int& m_blk, int& n_blk, int& kb0, int& kb1) {
    int split = w % p.num_splits;
    int tile = w / p.num_splits;
    m_blk = tile / num_n;
    n_blk = tile % num_n;
    kb0 = split * p.kb_per_split;
    kb1 = kb0 + p.kb_per_split;
    if (kb1 > num_kb_total) kb1 = num_kb_total;
  };

  if (warp == 0) {
    // ===================== TMA producer (one thread) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        int m_blk, n_blk, kb0, kb1;
        decode(w, m_blk, n_blk, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
          uint8_t* a_dst = smem_a + stage * S::kABytes;
          uint8_t* b_dst = smem_b + stage * S::kBBytes;
          if constexpr (A_MN) {
            // A stored [K, M]: boxes of [64 k][64 m]; chunk c covers m in [64c, 64c+64)
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)
              ptx::tma_load_2d(a_dst + c * (BLOCK_K * 128), &tmA, &full_bar[stage], m_blk * BLOCK_M + c * 64,
                               kb * BLOCK_K);
          } else {
            ptx::tma_load_2d(a_dst, &tmA, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              ptx::tma_load_2d(b_dst + c * (BLOCK_K * 128), &tmB, &full_bar[stage], n_blk * BN + c * 64,
                               kb * BLOCK_K);
          } else {
            ptx::tma_load_2d(b_dst, &tmB, &full_bar[stage], kb * BLOCK_K, n_blk * BN);
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
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(BLOCK_M, BN, true, A_MN, B_MN);
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
      for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
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
            ptx::umma_f16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        ptx::umma_commit(&tmem_full[buf]);  // accumulator ready for the epilogue
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps, 256 threads) =====================
    constexpr int kColsPerWarp = BN / 2;          // two warps share a lane quarter, each owns half the columns
    constexpr int kChunks = kColsPerWarp / 32;    // 32-column TMEM loads per warp
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;             // which half of the tile's columns
    const int row = q * 32 + lane;                // row of the tile owned by this thread
    const int epi_tid = threadIdx.x - 128;        // 0..255
    const bool stats_smem = (p.col_sum != nullptr) && (p.N <= kMaxStatCols);
    if (stats_smem) {
      for (int i = epi_tid; i < 2 * kMaxStatCols; i += kEpiThreads) colacc[i] = 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    int it = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
      int m_blk, n_blk, kb0, kb1;
      decode(w, m_blk, n_blk, kb0, kb1);
      const int buf = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      uint8_t* smem_d = smem_d0 + (S::kDBufs == 2 ? (it & 1) * S::kDBytes : 0);
      // the staging buffer about to be overwritten must have been drained by its previous TMA store
      if (epi_tid == 0) {
        if (S::kDBufs == 2) ptx::tma_wait_group_read<1>();
        else ptx::tma_wait_group_read<0>();
      }
      // residual addend: issue the global loads before blocking on the accumulator
      const int grow = m_blk * BLOCK_M + row;
      uint4 cadd[kChunks * 4];
      if (p.addend != nullptr) {
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          const int col0 = n_blk * BN + half * kColsPerWarp + c * 32;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (grow < p.M && col0 + i * 8 < p.N)
              cadd[c * 4 + i] = *reinterpret_cast<const uint4*>(p.addend + (int64_t)grow * p.ldc + col0 + i * 8);
            else
              cadd[c * 4 + i] = make_uint4(0, 0, 0, 0);
          }
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      ptx::mbar_wait(&tmem_full[buf], acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + buf * BN + half * kColsPerWarp + ((uint32_t)(q * 32) << 16);
      uint32_t r[kChunks][32];
#pragma unroll
      for (int c = 0; c < kChunks; ++c) ptx::tmem_ld_32x32b_x32(taddr + c * 32, r[c]);
      ptx::tmem_ld_wait();
      // all TMEM reads of this warp are done: hand the accumulator back to the MMA warp as early as possible
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[buf]);
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int cc = half * kChunks + c;              // 32-column chunk index within the tile
        const int col0 = n_blk * BN + cc * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(r[c][j]);
          if (p.bias != nullptr && (col0 + j) < p.N) x += __ldg(p.bias + col0 + j);
          v[j] = x;
        }
        if (p.addend != nullptr) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&cadd[c * 4 + i]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              v[i * 8 + 2 * k] += __low2float(h[k]);
              v[i * 8 + 2 * k + 1] += __high2float(h[k]);
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if constexpr (sizeof(OutT) == 2) {
          // 32 bf16 = 64B = 4 x 16B chunks; box = cc/2, logical 16B chunk = (cc%2)*4 + i, XOR-swizzled by row
          uint8_t* box = smem_d + (cc >> 1) * (BLOCK_M * 128) + row * 128;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 pk;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v[i * 8 + 0], v[i * 8 + 1]);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(v[i * 8 + 2], v[i * 8 + 3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[i * 8 + 4], v[i * 8 + 5]);
            __nv_bfloat162 h3 = __floats2bfloat162_rn(v[i * 8 + 6], v[i * 8 + 7]);
            pk.x = *reinterpret_cast<uint32_t*>(&h0);
            pk.y = *reinterpret_cast<uint32_t*>(&h1);
            pk.z = *reinterpret_cast<uint32_t*>(&h2);
            pk.w = *reinterpret_cast<uint32_t*>(&h3);
            const int chunk = ((cc & 1) * 4 + i) ^ (row & 7);
            *reinterpret_cast<uint4*>(box + chunk * 16) = pk;
          }
        } else {
          // 32 fp32 = 128B = one full swizzled row of box cc
          uint8_t* box = smem_d + cc * (BLOCK_M * 128) + row * 128;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            uint4 pk;
            pk.x = __float_as_uint(v[i * 4 + 0]);
            pk.y = __float_as_uint(v[i * 4 + 1]);
            pk.z = __float_as_uint(v[i * 4 + 2]);
            pk.w = __float_as_uint(v[i * 4 + 3]);
            const int chunk = i ^ (row & 7);
            *reinterpret_cast<uint4*>(box + chunk * 16) = pk;
          }
        }
      }
      // make generic-proxy smem writes visible to the async proxy, then one thread issues the TMA stores
      ptx::fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (p.col_sum != nullptr) {
        // per-column sum / sum-of-squares of the values as stored (rounded), for fused BatchNorm statistics.
        // thread t owns column t (t < BN); rows beyond M are zero because TMA zero-filled A.
        if constexpr (sizeof(OutT) == 2) {
          if (epi_tid < BN) {
            const int colg = n_blk * BN + epi_tid;
            if (colg < p.N) {
              const uint8_t* box = smem_d + (epi_tid >> 6) * (BLOCK_M * 128);
              const int cl = epi_tid & 63;
              float s = 0.f, s2 = 0.f;
#pragma unroll 8
              for (int rr = 0; rr < BLOCK_M; ++rr) {
                const int chunk = (cl >> 3) ^ (rr & 7);
                float x = __bfloat162float(
                    *reinterpret_cast<const __nv_bfloat16*>(box + rr * 128 + chunk * 16 + (cl & 7) * 2));
                s += x;
                s2 += x * x;
              }
              if (stats_smem) {   // one thread per column per tile: plain accumulate, flushed once at kernel end
                colacc[colg] += s;
                colacc[kMaxStatCols + colg] += s2;
              } else {
                atomicAdd(p.col_sum + colg, s);
                atomicAdd(p.col_sumsq + colg, s2);
              }
            }
          }
        }
      }
      if (epi_tid == 0) {
#pragma unroll
        for (int j = 0; j < kOutBoxes; ++j) {
          if constexpr (REDUCE)
            ptx::tma_reduce_add_2d(&tmD, smem_d + j * (BLOCK_M * 128), n_blk * BN + j * kOutBoxCols, m_blk * BLOCK_M);
          else
            ptx::tma_store_2d(&tmD, smem_d + j * (BLOCK_M * 128), n_blk * BN + j * kOutBoxCols, m_blk * BLOCK_M);
        }
        ptx::tma_commit_group();
      }
    }
    if (stats_smem) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int c = epi_tid; c < p.N; c += kEpiThreads) {
        const float a = colacc[c], b = colacc[kMaxStatCols + c];
        if (a != 0.f || b != 0.f) {
          atomicAdd(p.col_sum + c, a);
          atomicAdd(p.col_sumsq + c, b);
        }
      }
    }
    if (epi_tid == 0) ptx::tma_wait_group<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
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

template <int BN, bool A_MN, bool B_MN, typename OutT, bool REDUCE>
static void launch(const void* A, const void* B, void* D, int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldd,
                   const float* bias, int relu, float* col_sum, float* col_sumsq, const void* addend, int64_t ldc,
                   int splits, int max_ctas, const int* dbg, cudaStream_t stream) {
  using S = Smem<BN, OutT>;
  CUtensorMap tmA = A_MN ? make_map(A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, K, M, lda, BLOCK_K, 64)
                         : make_map(A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, K, lda, BLOCK_M, BLOCK_K);
  CUtensorMap tmB = B_MN ? make_map(B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, K, N, ldb, BLOCK_K, 64)
                         : make_map(B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, N, K, ldb, BN, BLOCK_K);
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
  p.dbg_lbo_a = dbg ? dbg[0] : 0; p.dbg_sbo_a = dbg ? dbg[1] : 0; p.dbg_lbo_b = dbg ? dbg[2] : 0; p.dbg_sbo_b = dbg ? dbg[3] : 0;
  const int num_work = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + BN - 1) / BN) * p.num_splits;
  int ctas = max_ctas > 0 ? max_ctas : kNumSMs;
  if (ctas > num_work) ctas = num_work;
  auto kern = gemm_kernel<BN, A_MN, B_MN, OutT, REDUCE>;
  CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
  kern<<<ctas, kThreads, S::kTotal, stream>>>(tmA, tmB, tmD, p);
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
  const bool bn64 = (block_n == 64) || (block_n == 0 && N <= 64);
  if (mode == 0) {
    if (bn64) gemm::launch<64, false, false, __nv_bfloat16, false>(A, B, D, M, N, K, lda, ldb, ldd, bias, relu, col_sum, col_sumsq, addend, ldc, 1, max_ctas, dbg, s);
    else gemm::launch<128, false, false, __nv_bfloat16, false>(A, B, D, M, N, K, lda, ldb, ldd, bias, relu, col_sum, col_sumsq, addend, ldc, 1, max_ctas, dbg, s);
  } else if (mode == 1) {
    if (bn64) gemm::launch<64, false, true, __nv_bfloat16, false>(A, B, D, M, N, K, lda, ldb, ldd, bias, relu, col_sum, col_sumsq, addend, ldc, 1, max_ctas, dbg, s);
    else gemm::launch<128, false, true, __nv_bfloat16, false>(A, B, D, M, N, K, lda, ldb, ldd, bias, relu, col_sum, col_sumsq, addend, ldc, 1, max_ctas, dbg, s);
  } else if (mode == 2) {
    if (bn64) gemm::launch<64, true, true, float, true>(A, B, D, M, N, K, lda, ldb, ldd, nullptr, 0, nullptr, nullptr, nullptr, 0, splits, max_ctas, dbg, s);
    else gemm::launch<128, true, true, float, true>(A, B, D, M, N, K, lda, ldb, ldd, nullptr, 0, nullptr, nullptr, nullptr, 0, splits, max_ctas, dbg, s);
  } else {
    throw std::runtime_error("MXBGemmBf16: bad mode");
  }
  MXB_CATCH
}
