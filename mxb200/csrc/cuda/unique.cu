// Sorted unique row ids for row_sparse push / pull (reference: src/kvstore/kvstore_utils.cu UniqueImpl = CUB radix sort +
// unique over the id list, O(n log n) with a temporary of the list's size).
//
// Row ids are bounded by the number of rows of the parameter, so sorting is unnecessary: (1) mark ids in a bitmap of
// num_rows bits, (2) count the set bits per 1024-row block, (3) one block scans the per-block counts, (4) every block
// writes its rows in order at its offset.  O(n + num_rows / 32) work, output sorted by construction, no atomics on the
// output.  The bitmap (num_rows / 8 bytes) and the per-block counters live in a caller-provided scratch buffer.
#include "common.cuh"

namespace mxb {
namespace uniq {

constexpr int kThreads = 256;
constexpr int kRowsPerBlock = 8192;                      // 256 words of 32 rows per block

__global__ void __launch_bounds__(kThreads) mark_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t num_rows, uint32_t* bitmap,
                                                       int* bad) {
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const int64_t r = ids[i];
    if (r < 0 || r >= num_rows) {
      *bad = 1;
      continue;
    }
    atomicOr(&bitmap[r >> 5], 1u << (r & 31));
  }
}

// one word per thread; block b covers words [256 b, 256 b + 256)
__global__ void __launch_bounds__(kThreads) count_kernel(const uint32_t* __restrict__ bitmap, int64_t nwords, int* block_count) {
  __shared__ int red[kThreads / 32];
  const int64_t w = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  int c = w < nwords ? __popc(bitmap[w]) : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int i = 0; i < kThreads / 32; ++i) s += red[i];
    block_count[blockIdx.x] = s;
  }
}

// exclusive scan of the per-block counts by ONE block (up to millions of blocks: chunks of 256 with a running carry)
__global__ void __launch_bounds__(kThreads) scan_kernel(int* block_count, int nblocks, int64_t* total) {
  __shared__ int buf[kThreads];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += kThreads) {
    const int i = base + threadIdx.x;
    const int v = i < nblocks ? block_count[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < kThreads; o <<= 1) {             // Hillis-Steele inclusive scan
      const int t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nblocks) block_count[i] = carry + buf[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == kThreads - 1) carry += buf[kThreads - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(kThreads) compact_kernel(uint32_t* bitmap, int64_t nwords, const int* __restrict__ block_off, int64_t* out) {
  __shared__ int buf[kThreads];
  const int64_t w = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const uint32_t bits = w < nwords ? bitmap[w] : 0u;
  if (w < nwords) bitmap[w] = 0u;                          // leave the scratch bitmap zeroed for the next call
  const int c = __popc(bits);
  buf[threadIdx.x] = c;
  __syncthreads();
  for (int o = 1; o < kThreads; o <<= 1) {
    const int t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
    __syncthreads();
    buf[threadIdx.x] += t;
    __syncthreads();
  }
  int64_t pos = (int64_t)block_off[blockIdx.x] + buf[threadIdx.x] - c;
  uint32_t b = bits;
  while (b) {
    const int bit = __ffs(b) - 1;
    out[pos++] = w * 32 + bit;
    b &= b - 1;
  }
}

}  // namespace uniq
}  // namespace mxb

using namespace mxb;

// bytes of scratch needed for ids in [0, num_rows)
MXB_API int MXBUniqueScratchBytes(int64_t num_rows, int64_t* bytes) {
  MXB_TRY
  const int64_t nwords = (num_rows + 31) / 32;
  const int64_t nblocks = (nwords + uniq::kThreads - 1) / uniq::kThreads;
  *bytes = nwords * 4 + nblocks * 4 + 64;
  MXB_CATCH
}

// ids: int64 [n] on the device; out: int64 [min(n, num_rows)]; scratch: zero-initialised on first use, left zeroed;
// count_dev: int64 [1] (number of unique ids); status_dev: int32 [1], set to 1 if an id was out of range
MXB_API int MXBUniqueRowIds(const void* ids, int64_t n, int64_t num_rows, void* out, void* scratch, void* count_dev, void* status_dev,
                            void* stream) {
  MXB_TRY
  MXB_CHECK(num_rows > 0 && num_rows < ((int64_t)1 << 40), "unique: bad row count");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t nwords = (num_rows + 31) / 32;
  const int64_t nblocks64 = (nwords + uniq::kThreads - 1) / uniq::kThreads;
  MXB_CHECK(nblocks64 < (1ll << 31), "unique: too many rows");
  const int nblocks = (int)nblocks64;
  uint32_t* bitmap = static_cast<uint32_t*>(scratch);
  int* block_count = reinterpret_cast<int*>(bitmap + nwords);
  if (n > 0) {
    const int grid = grid_for(n, uniq::kThreads, 4, 16);
    uniq::mark_kernel<<<grid, uniq::kThreads, 0, s>>>(static_cast<const int64_t*>(ids), n, num_rows, bitmap, static_cast<int*>(status_dev));
  }
  uniq::count_kernel<<<nblocks, uniq::kThreads, 0, s>>>(bitmap, nwords, block_count);
  uniq::scan_kernel<<<1, uniq::kThreads, 0, s>>>(block_count, nblocks, static_cast<int64_t*>(count_dev));
  uniq::compact_kernel<<<nblocks, uniq::kThreads, 0, s>>>(bitmap, nwords, block_count, static_cast<int64_t*>(out));
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}
