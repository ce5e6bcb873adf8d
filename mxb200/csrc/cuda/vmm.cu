// Symmetric, peer-mapped AND switch-multicast device memory for the fused collectives (NVLS path).
//
// cudaMalloc + CUDA-IPC (allreduce.cu: MXBIpc*) gives every rank a unicast mapping of every peer's arena -- enough for
// the P2P load/store kernels.  In-switch reduction (multimem.ld_reduce) and broadcast stores (multimem.st) additionally
// need ONE multicast object that all GPUs bind their physical allocation to, which only the driver's virtual memory
// management API provides:
//
//   cuMemCreate (per rank, POSIX-fd shareable)  -> physical arena
//   cuMulticastCreate (rank 0) / import (others) -> cuMulticastAddDevice (all) -> cuMulticastBindMem (all)
//   cuMemMap x (1 local + N-1 peers + 1 multicast) + cuMemSetAccess
//
// Handles cross process boundaries as file descriptors sent over abstract-namespace unix sockets (SCM_RIGHTS).
// Reference analogue: CommDevice::EnableP2P + InitMergeBuffer (src/kvstore/comm.h:678-755) and KVStoreNCCL's communicator
// bootstrap (src/kvstore/kvstore_nccl.h:90-140), i.e. "make every GPU's reduction buffer reachable from every GPU".
//
// All driver entry points are resolved through cudaGetDriverEntryPoint so that the library still loads on a host
// without libcuda (the CPU-only build / test box).
#include <errno.h>
#include <stdio.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace mxb {
namespace vmm {

#define DRV_FN(name, ...) typedef CUresult (*name##_t)(__VA_ARGS__)
DRV_FN(cuMemCreate, CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
DRV_FN(cuMemRelease, CUmemGenericAllocationHandle);
DRV_FN(cuMemGetAllocationGranularity, size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
DRV_FN(cuMemAddressReserve, CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
DRV_FN(cuMemAddressFree, CUdeviceptr, size_t);
DRV_FN(cuMemMap, CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
DRV_FN(cuMemUnmap, CUdeviceptr, size_t);
DRV_FN(cuMemSetAccess, CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
DRV_FN(cuMemExportToShareableHandle, void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
DRV_FN(cuMemImportFromShareableHandle, CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
DRV_FN(cuMulticastCreate, CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
DRV_FN(cuMulticastAddDevice, CUmemGenericAllocationHandle, CUdevice);
DRV_FN(cuMulticastBindMem, CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t,
       unsigned long long);
DRV_FN(cuMulticastGetGranularity, size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
DRV_FN(cuDeviceGetAttribute, int*, CUdevice_attribute, CUdevice);
DRV_FN(cuGetErrorString, CUresult, const char**);

struct Driver {
#define DRV_MEMBER(name) name##_t name = nullptr
  DRV_MEMBER(cuMemCreate);
  DRV_MEMBER(cuMemRelease);
  DRV_MEMBER(cuMemGetAllocationGranularity);
  DRV_MEMBER(cuMemAddressReserve);
  DRV_MEMBER(cuMemAddressFree);
  DRV_MEMBER(cuMemMap);
  DRV_MEMBER(cuMemUnmap);
  DRV_MEMBER(cuMemSetAccess);
  DRV_MEMBER(cuMemExportToShareableHandle);
  DRV_MEMBER(cuMemImportFromShareableHandle);
  DRV_MEMBER(cuMulticastCreate);
  DRV_MEMBER(cuMulticastAddDevice);
  DRV_MEMBER(cuMulticastBindMem);
  DRV_MEMBER(cuMulticastGetGranularity);
  DRV_MEMBER(cuDeviceGetAttribute);
  DRV_MEMBER(cuGetErrorString);
  bool ok = false;
};

static const Driver& drv() {
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    bool all = true;
    auto get = [&](const char* name, void** out) {
      cudaDriverEntryPointQueryResult q;
      void* f = nullptr;
      if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !f) {
        all = false;
        cudaGetLastError();
        return;
      }
      *out = f;
    };
#define DRV_GET(name) get(#name, reinterpret_cast<void**>(&d.name))
    DRV_GET(cuMemCreate);
    DRV_GET(cuMemRelease);
    DRV_GET(cuMemGetAllocationGranularity);
    DRV_GET(cuMemAddressReserve);
    DRV_GET(cuMemAddressFree);
    DRV_GET(cuMemMap);
    DRV_GET(cuMemUnmap);
    DRV_GET(cuMemSetAccess);
    DRV_GET(cuMemExportToShareableHandle);
    DRV_GET(cuMemImportFromShareableHandle);
    DRV_GET(cuMulticastCreate);
    DRV_GET(cuMulticastAddDevice);
    DRV_GET(cuMulticastBindMem);
    DRV_GET(cuMulticastGetGranularity);
    DRV_GET(cuDeviceGetAttribute);
    DRV_GET(cuGetErrorString);
    d.ok = all;
  });
  return d;
}

static void cu_check(CUresult r, const char* what) {
  if (r == CUDA_SUCCESS) return;
  const char* msg = nullptr;
  if (drv().cuGetErrorString) drv().cuGetErrorString(r, &msg);
  throw std::runtime_error(std::string("driver call failed: ") + what + ": " + (msg ? msg : "?") + " (" + std::to_string((int)r) + ")");
}
#define CU_CHECK(expr) cu_check((expr), #expr)

// One symmetric allocation as seen by this rank.
struct Segment {
  int dev = 0, rank = 0, world = 1;
  size_t bytes = 0;                               // padded size (multiple of every granularity involved)
  CUmemGenericAllocationHandle phys = 0;          // my physical memory
  CUmemGenericAllocationHandle mc = 0;            // the multicast object (0: no multicast)
  std::vector<CUmemGenericAllocationHandle> peer; // imported physical handles (0 at my own index)
  std::vector<CUdeviceptr> va;                    // unicast mapping per rank (va[rank] = mine)
  CUdeviceptr mc_va = 0;
  int my_fd = -1, mc_fd = -1;
};

static CUmemAllocationProp phys_prop(int dev) {
  CUmemAllocationProp p;
  memset(&p, 0, sizeof(p));
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = dev;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

static void map_rw(CUdeviceptr va, size_t bytes, int dev) {
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = dev;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CU_CHECK(drv().cuMemSetAccess(va, bytes, &acc, 1));
}

// ---------------------------------------------------------------- fd passing over abstract unix sockets
static std::string sock_name(const char* job, int tag, int rank) {
  return std::string("mxb200-vmm-") + job + "-" + std::to_string(tag) + "-" + std::to_string(rank);
}

static int listen_abstract(const std::string& name) {
  int s = socket(AF_UNIX, SOCK_STREAM, 0);
  if (s < 0) throw std::runtime_error("vmm: socket() failed");
  sockaddr_un a;
  memset(&a, 0, sizeof(a));
  a.sun_family = AF_UNIX;
  if (name.size() + 2 > sizeof(a.sun_path)) throw std::runtime_error("vmm: socket name too long");
  memcpy(a.sun_path + 1, name.data(), name.size());     // leading NUL: abstract namespace, nothing to unlink
  if (bind(s, reinterpret_cast<sockaddr*>(&a), (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + name.size())) != 0) {
    close(s);
    throw std::runtime_error("vmm: bind(" + name + ") failed: " + strerror(errno));
  }
  if (listen(s, 64) != 0) {
    close(s);
    throw std::runtime_error("vmm: listen failed");
  }
  return s;
}

static void send_fds(const std::string& name, int from_rank, const int* fds, int nfds) {
  int s = socket(AF_UNIX, SOCK_STREAM, 0);
  if (s < 0) throw std::runtime_error("vmm: socket() failed");
  sockaddr_un a;
  memset(&a, 0, sizeof(a));
  a.sun_family = AF_UNIX;
  memcpy(a.sun_path + 1, name.data(), name.size());
  const socklen_t len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + name.size());
  int tries = 0;
  while (connect(s, reinterpret_cast<sockaddr*>(&a), len) != 0) {
    if (++tries > 3000) {
      close(s);
      throw std::runtime_error("vmm: connect(" + name + ") failed: " + strerror(errno));
    }
    usleep(10000);
  }
  int payload[2] = {from_rank, nfds};
  iovec iov = {payload, sizeof(payload)};
  char ctrl[CMSG_SPACE(sizeof(int) * 4)];
  memset(ctrl, 0, sizeof(ctrl));
  msghdr msg;
  memset(&msg, 0, sizeof(msg));
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = CMSG_SPACE(sizeof(int) * nfds);
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET;
  c->cmsg_type = SCM_RIGHTS;
  c->cmsg_len = CMSG_LEN(sizeof(int) * nfds);
  memcpy(CMSG_DATA(c), fds, sizeof(int) * nfds);
  if (sendmsg(s, &msg, 0) < 0) {
    close(s);
    throw std::runtime_error(std::string("vmm: sendmsg failed: ") + strerror(errno));
  }
  close(s);
}

static void recv_fds(int lsock, int* from_rank, int* fds, int* nfds) {
  int c = accept(lsock, nullptr, nullptr);
  if (c < 0) throw std::runtime_error(std::string("vmm: accept failed: ") + strerror(errno));
  int payload[2] = {0, 0};
  iovec iov = {payload, sizeof(payload)};
  char ctrl[CMSG_SPACE(sizeof(int) * 4)];
  memset(ctrl, 0, sizeof(ctrl));
  msghdr msg;
  memset(&msg, 0, sizeof(msg));
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  if (recvmsg(c, &msg, MSG_WAITALL) < 0) {
    close(c);
    throw std::runtime_error(std::string("vmm: recvmsg failed: ") + strerror(errno));
  }
  close(c);
  *from_rank = payload[0];
  *nfds = payload[1];
  cmsghdr* h = CMSG_FIRSTHDR(&msg);
  if (!h || h->cmsg_type != SCM_RIGHTS) throw std::runtime_error("vmm: no file descriptors in message");
  memcpy(fds, CMSG_DATA(h), sizeof(int) * (*nfds));
}

}  // namespace vmm
}  // namespace mxb

using namespace mxb;
using namespace mxb::vmm;

// 1 if this device (and driver) can create multicast objects.
MXB_API int MXBVmmMulticastSupported(int dev, int* out) {
  MXB_TRY
  *out = 0;
  if (!drv().ok) return 0;
  int v = 0;
  if (drv().cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, (CUdevice)dev) == CUDA_SUCCESS) *out = v;
  MXB_CATCH
}

// Phase 1 (every rank): create the local physical arena, map it, export its fd, open the rendezvous socket.
//   world == 1 or ``multicast`` == 0 skips the multicast object.  Returns an opaque segment handle and the padded size.
MXB_API int MXBVmmCreate(int dev, int rank, int world, int64_t bytes, int multicast, void** seg_out, int64_t* padded_out,
                         void** local_ptr) {
  MXB_TRY
  MXB_CHECK(drv().ok, "CUDA driver VMM entry points unavailable");
  CUDA_CHECK(cudaSetDevice(dev));
  CUDA_CHECK(cudaFree(0));
  auto* sg = new Segment();
  sg->dev = dev; sg->rank = rank; sg->world = world;
  CUmemAllocationProp prop = phys_prop(dev);
  size_t gran = 0;
  CU_CHECK(drv().cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  if (multicast && world > 1) {
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof(mp));
    mp.numDevices = (unsigned)world;
    mp.size = (size_t)bytes;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mg = 0;
    CU_CHECK(drv().cuMulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    if (mg > gran) gran = mg;
  }
  sg->bytes = ((size_t)bytes + gran - 1) / gran * gran;
  CU_CHECK(drv().cuMemCreate(&sg->phys, sg->bytes, &prop, 0));
  sg->va.assign(world, 0);
  sg->peer.assign(world, 0);
  CU_CHECK(drv().cuMemAddressReserve(&sg->va[rank], sg->bytes, gran, 0, 0));
  CU_CHECK(drv().cuMemMap(sg->va[rank], sg->bytes, 0, sg->phys, 0));
  map_rw(sg->va[rank], sg->bytes, dev);
  CUDA_CHECK(cudaMemset(reinterpret_cast<void*>(sg->va[rank]), 0, sg->bytes));
  CUDA_CHECK(cudaDeviceSynchronize());
  if (world > 1) CU_CHECK(drv().cuMemExportToShareableHandle(&sg->my_fd, sg->phys, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  *seg_out = sg;
  *padded_out = (int64_t)sg->bytes;
  *local_ptr = reinterpret_cast<void*>(sg->va[rank]);
  MXB_CATCH
}

// Phase 2 (every rank, collective; the caller must have passed a barrier after phase 1 so that every listener exists
// -- we bind here and retry connects, so only "everybody eventually calls this" is required):
// exchange physical-memory fds all-to-all (+ the multicast fd from rank 0), map the peers.
MXB_API int MXBVmmExchange(void* seg, const char* job, int tag, int multicast) {
  MXB_TRY
  auto* sg = reinterpret_cast<Segment*>(seg);
  if (sg->world == 1) return 0;
  CUDA_CHECK(cudaSetDevice(sg->dev));
  const bool mc = multicast != 0;
  if (mc && sg->rank == 0) {
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof(mp));
    mp.numDevices = (unsigned)sg->world;
    mp.size = sg->bytes;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CU_CHECK(drv().cuMulticastCreate(&sg->mc, &mp));
    CU_CHECK(drv().cuMemExportToShareableHandle(&sg->mc_fd, sg->mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  }
  const int lsock = listen_abstract(sock_name(job, tag, sg->rank));
  try {
    for (int k = 1; k < sg->world; ++k) {
      const int to = (sg->rank + k) % sg->world;
      int fds[2] = {sg->my_fd, sg->mc_fd};
      send_fds(sock_name(job, tag, to), sg->rank, fds, (mc && sg->rank == 0) ? 2 : 1);
    }
    for (int k = 1; k < sg->world; ++k) {
      int from = -1, fds[4] = {-1, -1, -1, -1}, n = 0;
      recv_fds(lsock, &from, fds, &n);
      MXB_CHECK(from >= 0 && from < sg->world && from != sg->rank && n >= 1, "vmm: bad fd message");
      CU_CHECK(drv().cuMemImportFromShareableHandle(&sg->peer[from], reinterpret_cast<void*>((uintptr_t)fds[0]),
                                                    CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
      close(fds[0]);
      if (n >= 2) {
        MXB_CHECK(from == 0 && mc, "vmm: unexpected multicast handle");
        CU_CHECK(drv().cuMemImportFromShareableHandle(&sg->mc, reinterpret_cast<void*>((uintptr_t)fds[1]),
                                                      CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
        close(fds[1]);
      }
    }
  } catch (...) {
    close(lsock);
    throw;
  }
  close(lsock);
  CUmemAllocationProp prop = phys_prop(sg->dev);
  size_t gran = 0;
  CU_CHECK(drv().cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  for (int r = 0; r < sg->world; ++r) {
    if (r == sg->rank) continue;
    CU_CHECK(drv().cuMemAddressReserve(&sg->va[r], sg->bytes, gran, 0, 0));
    CU_CHECK(drv().cuMemMap(sg->va[r], sg->bytes, 0, sg->peer[r], 0));
    map_rw(sg->va[r], sg->bytes, sg->dev);
  }
  if (mc) CU_CHECK(drv().cuMulticastAddDevice(sg->mc, (CUdevice)sg->dev));
  MXB_CATCH
}

// Phase 3 (every rank, after a barrier: all devices must have been added before anyone binds): bind + map multicast.
MXB_API int MXBVmmBindMulticast(void* seg) {
  MXB_TRY
  auto* sg = reinterpret_cast<Segment*>(seg);
  if (sg->world == 1 || sg->mc == 0) return 0;
  CUDA_CHECK(cudaSetDevice(sg->dev));
  CU_CHECK(drv().cuMulticastBindMem(sg->mc, 0, sg->phys, 0, sg->bytes, 0));
  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof(mp));
  mp.numDevices = (unsigned)sg->world;
  mp.size = sg->bytes;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t mg = 0;
  CU_CHECK(drv().cuMulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
  CU_CHECK(drv().cuMemAddressReserve(&sg->mc_va, sg->bytes, mg, 0, 0));
  CU_CHECK(drv().cuMemMap(sg->mc_va, sg->bytes, 0, sg->mc, 0));
  map_rw(sg->mc_va, sg->bytes, sg->dev);
  MXB_CATCH
}

// Single-process variant (one process drives all GPUs): segments share handles directly, no fds.
MXB_API int MXBVmmLinkLocal(void** segs, int world, int multicast) {
  MXB_TRY
  std::vector<Segment*> sg(world);
  for (int r = 0; r < world; ++r) sg[r] = reinterpret_cast<Segment*>(segs[r]);
  if (world == 1) return 0;
  CUmemGenericAllocationHandle mc = 0;
  if (multicast) {
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof(mp));
    mp.numDevices = (unsigned)world;
    mp.size = sg[0]->bytes;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CU_CHECK(drv().cuMulticastCreate(&mc, &mp));
    for (int r = 0; r < world; ++r) CU_CHECK(drv().cuMulticastAddDevice(mc, (CUdevice)sg[r]->dev));
  }
  for (int r = 0; r < world; ++r) {
    CUDA_CHECK(cudaSetDevice(sg[r]->dev));
    CUmemAllocationProp prop = phys_prop(sg[r]->dev);
    size_t gran = 0;
    CU_CHECK(drv().cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    for (int q = 0; q < world; ++q) {
      if (q == r) continue;
      CU_CHECK(drv().cuMemAddressReserve(&sg[r]->va[q], sg[r]->bytes, gran, 0, 0));
      CU_CHECK(drv().cuMemMap(sg[r]->va[q], sg[r]->bytes, 0, sg[q]->phys, 0));
      map_rw(sg[r]->va[q], sg[r]->bytes, sg[r]->dev);
    }
    sg[r]->mc = mc;
  }
  if (multicast)
    for (int r = 0; r < world; ++r) {
      auto* s = sg[r];
      MXBVmmBindMulticast(s) == 0 ? void() : throw std::runtime_error(MXBGetLastError());
    }
  MXB_CATCH
}

// ptrs_out[world] = unicast mapping of every rank's arena in this process; *mc_out = multicast mapping or null.
MXB_API int MXBVmmPointers(void* seg, void** ptrs_out, void** mc_out) {
  MXB_TRY
  auto* sg = reinterpret_cast<Segment*>(seg);
  for (int r = 0; r < sg->world; ++r) ptrs_out[r] = reinterpret_cast<void*>(sg->va[r]);
  *mc_out = reinterpret_cast<void*>(sg->mc_va);
  MXB_CATCH
}

MXB_API int MXBVmmFree(void* seg) {
  MXB_TRY
  auto* sg = reinterpret_cast<Segment*>(seg);
  if (!sg) return 0;
  cudaSetDevice(sg->dev);
  cudaDeviceSynchronize();
  if (sg->mc_va) {
    drv().cuMemUnmap(sg->mc_va, sg->bytes);
    drv().cuMemAddressFree(sg->mc_va, sg->bytes);
  }
  for (int r = 0; r < sg->world; ++r) {
    if (sg->va[r]) {
      drv().cuMemUnmap(sg->va[r], sg->bytes);
      drv().cuMemAddressFree(sg->va[r], sg->bytes);
    }
    if (sg->peer[r]) drv().cuMemRelease(sg->peer[r]);
  }
  if (sg->my_fd >= 0) close(sg->my_fd);
  if (sg->mc_fd >= 0) close(sg->mc_fd);
  if (sg->mc && sg->rank == 0) drv().cuMemRelease(sg->mc);
  else if (sg->mc && sg->world > 1 && sg->my_fd >= 0) drv().cuMemRelease(sg->mc);   // imported handle (process mode)
  if (sg->phys) drv().cuMemRelease(sg->phys);
  delete sg;
  MXB_CATCH
}
