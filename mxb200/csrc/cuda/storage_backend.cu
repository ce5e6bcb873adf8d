// Raw device / pinned-host allocation back-ends of the pooled storage manager (csrc/host/storage.cc): the host library
// has no CUDA dependency, python hands these entry points to MXBStorageSetBackend.
// Reference: src/storage/gpu_device_storage.h (cudaMalloc / cudaFree), pinned_memory_storage.h (cudaHostAlloc portable).
#include "common.cuh"

namespace {
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (dev >= 0 && dev != prev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
}  // namespace

MXB_API int MXBRawDeviceAlloc(int dev, int64_t bytes, void** out) {
  DeviceGuard g(dev);
  *out = nullptr;
  const cudaError_t e = cudaMalloc(out, (size_t)bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();          // clear the sticky out-of-memory so the manager can release its pool and retry
    return 1;
  }
  return 0;
}

MXB_API int MXBRawDeviceFree(int dev, void* p) {
  DeviceGuard g(dev);
  return cudaFree(p) == cudaSuccess ? 0 : 1;
}

MXB_API int MXBRawPinnedAlloc(int dev, int64_t bytes, void** out) {
  DeviceGuard g(dev);
  *out = nullptr;
  const cudaError_t e = cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 1;
  }
  return 0;
}

MXB_API int MXBRawPinnedFree(int dev, void* p) {
  (void)dev;
  return cudaFreeHost(p) == cudaSuccess ? 0 : 1;
}

MXB_API int MXBRawMemInfo(int dev, int64_t* free_bytes, int64_t* total_bytes) {
  DeviceGuard g(dev);
  size_t f = 0, t = 0;
  if (cudaMemGetInfo(&f, &t) != cudaSuccess) {
    cudaGetLastError();
    return 1;
  }
  *free_bytes = (int64_t)f;
  *total_bytes = (int64_t)t;
  return 0;
}
