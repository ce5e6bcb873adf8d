// Pooled storage manager (reference: include/mxnet/storage.h, src/storage/storage.cc,
// src/storage/pooled_storage_manager.h: GPUPooledStorageManager "Naive" and GPUPooledRoundedStorageManager "Round",
// src/storage/pinned_memory_storage.h, cpu_device_storage.h).
//
// One manager per (device type, device id).  Alloc rounds the request to a bucket, pops a block of that bucket if the pool
// has one and otherwise asks the device back-end; Free returns the block to its bucket; DirectFree hands it back to the
// back-end; ReleaseAll empties the pool.  The reserve rule of the reference is kept: if a raw allocation would leave less
// than MXNET_GPU_MEM_POOL_RESERVE percent of the device free, the pool is released first.
//
//   bucket("Naive")  = size rounded up to the page size (MXNET_GPU_MEM_POOL_PAGE_SIZE, 4096)
//   bucket("Round")  = multiples of the page below 2^cutoff (MXNET_GPU_MEM_POOL_ROUND_LINEAR_CUTOFF, 24), powers of two above
//
// The raw back-ends of device / pinned memory live in the CUDA library (this library has no CUDA dependency so that it
// loads on any host): python passes their entry points to MXBStorageSetBackend.  Plain host memory is served here
// (posix_memalign, 64-byte aligned: cache lines / AVX-512; the reference's CPUDeviceStorage aligns to 16).
// What this pool is for in this framework: the symmetric gradient / weight arenas (so an elastic re-init re-uses the
// blocks of the previous membership instead of cudaFree + cudaMalloc), pinned staging of the input pipeline and of
// checkpoint snapshots, kernel workspaces.  torch's caching allocator keeps serving torch tensors.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../c_api.h"

namespace {

enum DevType { kCPU = 1, kGPU = 2, kPinned = 3 };

typedef int (*RawAlloc)(int dev_id, int64_t bytes, void** out);
typedef int (*RawFree)(int dev_id, void* p);
typedef int (*RawMemInfo)(int dev_id, int64_t* free_bytes, int64_t* total_bytes);

struct Backend {
  RawAlloc alloc = nullptr;
  RawFree free = nullptr;
  RawMemInfo info = nullptr;
};

struct Manager {
  std::mutex mu;
  std::map<int64_t, std::vector<void*>> pool;              // bucket size -> free blocks
  std::unordered_map<void*, int64_t> live;                 // handed-out block -> bucket size
  int64_t used = 0, pooled = 0, n_alloc = 0, n_raw = 0, n_release = 0;
};

struct Storage {
  std::mutex mu;
  std::map<std::pair<int, int>, Manager*> managers;
  Backend backends[4];
  bool round = false;
  int64_t page = 4096;
  int cutoff = 24;
  int reserve = 5;

  Storage() {
    const char* t = getenv("MXNET_GPU_MEM_POOL_TYPE");
    round = t != nullptr && std::string(t) == "Round";
    if (const char* e = getenv("MXNET_GPU_MEM_POOL_PAGE_SIZE")) page = std::max<int64_t>(64, atoll(e));
    if (const char* e = getenv("MXNET_GPU_MEM_POOL_ROUND_LINEAR_CUTOFF")) cutoff = std::min(40, std::max(12, atoi(e)));
    if (const char* e = getenv("MXNET_GPU_MEM_POOL_RESERVE")) reserve = std::min(100, std::max(0, atoi(e)));
  }

  Manager* manager(int dev_type, int dev_id) {
    std::lock_guard<std::mutex> lk(mu);
    Manager*& m = managers[{dev_type, dev_id}];
    if (m == nullptr) m = new Manager();
    return m;
  }

  int64_t bucket(int64_t size) const {
    if (size <= 0) size = 1;
    if (round && size > ((int64_t)1 << cutoff)) {
      int64_t b = (int64_t)1 << cutoff;
      while (b < size) b <<= 1;
      return b;
    }
    return (size + page - 1) / page * page;
  }

  void raw_free(int dev_type, int dev_id, void* p) {
    if (dev_type == kCPU) {
      ::free(p);
      return;
    }
    MXB_CHECK(backends[dev_type].free != nullptr, "storage: no back-end registered for this device type");
    if (backends[dev_type].free(dev_id, p) != 0) throw std::runtime_error("storage: device free failed");
  }

  void release_all(int dev_type, int dev_id, Manager* m) {      // m->mu held
    for (auto& kv : m->pool)
      for (void* p : kv.second) raw_free(dev_type, dev_id, p);
    m->pool.clear();
    m->pooled = 0;
    m->n_release++;
  }

  void* raw_alloc(int dev_type, int dev_id, int64_t bytes, Manager* m) {   // m->mu held
    if (dev_type == kCPU) {
      void* p = nullptr;
      if (posix_memalign(&p, 64, (size_t)bytes) != 0) throw std::runtime_error("storage: host allocation failed");
      return p;
    }
    const Backend& b = backends[dev_type];
    MXB_CHECK(b.alloc != nullptr, "storage: no back-end registered for this device type (is the CUDA library loaded?)");
    if (dev_type == kGPU && b.info != nullptr && reserve > 0) {
      int64_t fr = 0, tot = 0;
      if (b.info(dev_id, &fr, &tot) == 0 && tot > 0 && (fr - bytes) < tot / 100 * reserve && m->pooled > 0)
        release_all(dev_type, dev_id, m);                      // pooled_storage_manager.h:108-111
    }
    void* p = nullptr;
    if (b.alloc(dev_id, bytes, &p) != 0 || p == nullptr) {
      release_all(dev_type, dev_id, m);                        // out of memory: give the pool back and retry once
      if (b.alloc(dev_id, bytes, &p) != 0 || p == nullptr)
        throw std::runtime_error("storage: device allocation of " + std::to_string(bytes) + " bytes failed");
    }
    return p;
  }
};

Storage* storage() {
  static Storage* s = new Storage();      // never destroyed: blocks may outlive static destruction order
  return s;
}

}  // namespace

MXB_API int MXBStorageSetBackend(int dev_type, void* alloc_fn, void* free_fn, void* meminfo_fn) {
  MXB_TRY
  MXB_CHECK(dev_type == kGPU || dev_type == kPinned, "storage back-ends exist for gpu (2) and cpu_pinned (3)");
  Storage* s = storage();
  std::lock_guard<std::mutex> lk(s->mu);
  s->backends[dev_type].alloc = reinterpret_cast<RawAlloc>(alloc_fn);
  s->backends[dev_type].free = reinterpret_cast<RawFree>(free_fn);
  s->backends[dev_type].info = reinterpret_cast<RawMemInfo>(meminfo_fn);
  MXB_CATCH
}

// *bucket_bytes: the size actually reserved (>= bytes)
MXB_API int MXBStorageAlloc(int dev_type, int dev_id, int64_t bytes, void** out, int64_t* bucket_bytes) {
  MXB_TRY
  MXB_CHECK(dev_type >= kCPU && dev_type <= kPinned, "storage: unknown device type");
  Storage* s = storage();
  Manager* m = s->manager(dev_type, dev_id);
  const int64_t b = s->bucket(bytes);
  std::lock_guard<std::mutex> lk(m->mu);
  m->n_alloc++;
  void* p = nullptr;
  auto it = m->pool.find(b);
  if (it != m->pool.end() && !it->second.empty()) {
    p = it->second.back();
    it->second.pop_back();
    m->pooled -= b;
  } else {
    p = s->raw_alloc(dev_type, dev_id, b, m);
    m->n_raw++;
  }
  m->live[p] = b;
  m->used += b;
  *out = p;
  if (bucket_bytes) *bucket_bytes = b;
  MXB_CATCH
}

static int free_impl(int dev_type, int dev_id, void* p, bool direct) {
  MXB_TRY
  if (p == nullptr) return 0;
  Storage* s = storage();
  Manager* m = s->manager(dev_type, dev_id);
  std::lock_guard<std::mutex> lk(m->mu);
  auto it = m->live.find(p);
  MXB_CHECK(it != m->live.end(), "storage: pointer was not allocated by this manager (double free?)");
  const int64_t b = it->second;
  m->live.erase(it);
  m->used -= b;
  if (direct) {
    s->raw_free(dev_type, dev_id, p);
  } else {
    m->pool[b].push_back(p);
    m->pooled += b;
  }
  MXB_CATCH
}

MXB_API int MXBStorageFree(int dev_type, int dev_id, void* p) { return free_impl(dev_type, dev_id, p, false); }
MXB_API int MXBStorageDirectFree(int dev_type, int dev_id, void* p) { return free_impl(dev_type, dev_id, p, true); }

MXB_API int MXBStorageReleaseAll(int dev_type, int dev_id) {
  MXB_TRY
  Storage* s = storage();
  Manager* m = s->manager(dev_type, dev_id);
  std::lock_guard<std::mutex> lk(m->mu);
  s->release_all(dev_type, dev_id, m);
  MXB_CATCH
}

// stats[6] = bytes handed out, bytes cached in the pool, Alloc calls, raw back-end allocations, ReleaseAll count, live blocks
MXB_API int MXBStorageStats(int dev_type, int dev_id, int64_t* stats) {
  MXB_TRY
  Manager* m = storage()->manager(dev_type, dev_id);
  std::lock_guard<std::mutex> lk(m->mu);
  stats[0] = m->used;
  stats[1] = m->pooled;
  stats[2] = m->n_alloc;
  stats[3] = m->n_raw;
  stats[4] = m->n_release;
  stats[5] = (int64_t)m->live.size();
  MXB_CATCH
}

MXB_API int MXBStorageBucket(int64_t bytes, int64_t* bucket_bytes) {
  MXB_TRY
  *bucket_bytes = storage()->bucket(bytes);
  MXB_CATCH
}
