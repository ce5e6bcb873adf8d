// GPU half of the image input pipeline: batched JPEG decode (nvJPEG, loaded with dlopen so the kernel library has no
// link-time dependency on it) + ONE kernel that does everything the reference's CPU augmenter chain does per image
// (src/io/image_aug_default.cc: resize / random-resized crop / mirror; src/io/iter_normalize.h: mean / std) and writes
// the network's input layout and dtype directly (NHWC bf16 for our models, NCHW fp32 for the generic API).
//
// The reference decodes with OpenCV on `preprocess_threads` CPU threads (src/io/iter_image_recordio_2.cc:300-420) and
// ships fp32 NCHW to the GPU: 602 KB per 224x224 image over PCIe.  Here the COMPRESSED bytes (~110 KB per ImageNet
// image) cross the host link, the decode runs on the GPU and the decoded pixels are read exactly once by the
// augmentation kernel.
#include <dlfcn.h>
#include <nvjpeg.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace mxb {
namespace jpeg {

struct Api {
  void* lib = nullptr;
  decltype(&nvjpegCreateEx) CreateEx = nullptr;
  decltype(&nvjpegDestroy) Destroy = nullptr;
  decltype(&nvjpegJpegStateCreate) StateCreate = nullptr;
  decltype(&nvjpegJpegStateDestroy) StateDestroy = nullptr;
  decltype(&nvjpegDecodeBatchedInitialize) BatchedInit = nullptr;
  decltype(&nvjpegDecodeBatched) DecodeBatched = nullptr;
  decltype(&nvjpegDecode) Decode = nullptr;
};

static Api* api() {
  static Api a;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12",
                           "/usr/local/cuda/lib64/libnvjpeg.so"};
    for (const char* n : names) {
      a.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (a.lib) break;
    }
    if (!a.lib) {
      err = std::string("cannot load libnvjpeg: ") + dlerror();
      return;
    }
#define MXB_SYM(field, name)                                            \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, #name)); \
  if (!a.field) err = "libnvjpeg lacks " #name;
    MXB_SYM(CreateEx, nvjpegCreateEx)
    MXB_SYM(Destroy, nvjpegDestroy)
    MXB_SYM(StateCreate, nvjpegJpegStateCreate)
    MXB_SYM(StateDestroy, nvjpegJpegStateDestroy)
    MXB_SYM(BatchedInit, nvjpegDecodeBatchedInitialize)
    MXB_SYM(DecodeBatched, nvjpegDecodeBatched)
    MXB_SYM(Decode, nvjpegDecode)
#undef MXB_SYM
  });
  if (!err.empty()) throw std::runtime_error(err);
  return &a;
}

struct Decoder {
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
  int batch = 0;
  int threads = 1;
};

static void check(nvjpegStatus_t st, const char* what) {
  if (st != NVJPEG_STATUS_SUCCESS) throw std::runtime_error(std::string("nvjpeg ") + what + " failed: status " + std::to_string((int)st));
}

// One image of the batch: where its decoded RGB-interleaved pixels live and which window of it becomes the output.
struct Window {
  int64_t src_off;      // byte offset of pixel (0, 0) inside the decoded-pixel arena (row pitch = 3 * w)
  int w, h;             // decoded size
  float x0, y0, cw, ch; // crop window in source pixels (top-left corner, extent); resized to the output size
  int mirror;           // 1: flip horizontally
  int pad;
};
static_assert(sizeof(Window) == 40, "Window is mirrored by a numpy structured dtype in io/gpu_pipeline.py");

constexpr int kThreads = 256;

// out[n, oy, ox, c] (NHWC) or out[n, c, oy, ox] (NCHW) = (bilinear(src window) - mean[c]) * istd[c]
template <typename OutT, bool NCHW>
__global__ void __launch_bounds__(kThreads)
crop_resize_kernel(const uint8_t* __restrict__ arena, const Window* __restrict__ win, OutT* __restrict__ out, int n, int OH, int OW,
                   float m0, float m1, float m2, float s0, float s1, float s2) {
  const int img = blockIdx.y;
  const Window wd = win[img];
  const uint8_t* __restrict__ src = arena + wd.src_off;
  const float sx = wd.cw / (float)OW, sy = wd.ch / (float)OH;
  const int64_t pitch = 3 * (int64_t)wd.w;
  for (int pix = blockIdx.x * kThreads + threadIdx.x; pix < OH * OW; pix += gridDim.x * kThreads) {
    const int oy = pix / OW, ox0 = pix - oy * OW;
    const int ox = wd.mirror ? OW - 1 - ox0 : ox0;
    // pixel-centre convention of cv::resize(INTER_LINEAR)
    float fx = wd.x0 + ((float)ox + 0.5f) * sx - 0.5f, fy = wd.y0 + ((float)oy + 0.5f) * sy - 0.5f;
    fx = fminf(fmaxf(fx, 0.f), (float)(wd.w - 1));
    fy = fminf(fmaxf(fy, 0.f), (float)(wd.h - 1));
    const int x0 = (int)fx, y0 = (int)fy;
    const int x1 = min(x0 + 1, wd.w - 1), y1 = min(y0 + 1, wd.h - 1);
    const float ax = fx - (float)x0, ay = fy - (float)y0;
    const uint8_t* r0 = src + y0 * pitch;
    const uint8_t* r1 = src + y1 * pitch;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float a = (float)r0[3 * x0 + c], b = (float)r0[3 * x1 + c], d = (float)r1[3 * x0 + c], e = (float)r1[3 * x1 + c];
      const float top = a + (b - a) * ax, bot = d + (e - d) * ax;
      v[c] = top + (bot - top) * ay;
    }
    v[0] = (v[0] - m0) * s0;
    v[1] = (v[1] - m1) * s1;
    v[2] = (v[2] - m2) * s2;
    if constexpr (NCHW) {
      OutT* o = out + ((int64_t)img * 3) * OH * OW + (int64_t)oy * OW + ox0;
      o[0] = from_f<OutT>(v[0]);
      o[(int64_t)OH * OW] = from_f<OutT>(v[1]);
      o[2 * (int64_t)OH * OW] = from_f<OutT>(v[2]);
    } else {
      OutT* o = out + (((int64_t)img * OH + oy) * OW + ox0) * 3;
      o[0] = from_f<OutT>(v[0]);
      o[1] = from_f<OutT>(v[1]);
      o[2] = from_f<OutT>(v[2]);
    }
  }
}

}  // namespace jpeg
}  // namespace mxb

using namespace mxb;

MXB_API int MXBJpegAvailable(int* ok) {
  MXB_TRY
  *ok = 0;
  try {
    jpeg::api();
    *ok = 1;
  } catch (...) {
  }
  MXB_CATCH
}

// backend: nvjpegBackend_t (0 = default / GPU-hybrid, 3 = hardware engine); threads: host threads nvJPEG may use per batch
MXB_API int MXBJpegCreate(int backend, int threads, void** out) {
  MXB_TRY
  auto* a = jpeg::api();
  auto* d = new jpeg::Decoder();
  d->threads = threads > 0 ? threads : 1;
  try {
    jpeg::check(a->CreateEx((nvjpegBackend_t)backend, nullptr, nullptr, NVJPEG_FLAGS_DEFAULT, &d->handle), "create");
    jpeg::check(a->StateCreate(d->handle, &d->state), "state create");
  } catch (...) {
    if (d->handle) a->Destroy(d->handle);
    delete d;
    throw;
  }
  *out = d;
  MXB_CATCH
}

MXB_API int MXBJpegFree(void* h) {
  MXB_TRY
  auto* a = jpeg::api();
  auto* d = static_cast<jpeg::Decoder*>(h);
  if (d->state) a->StateDestroy(d->state);
  if (d->handle) a->Destroy(d->handle);
  delete d;
  MXB_CATCH
}

// Decode n JPEG streams (host memory: blob + offsets[i], lengths[i]) to RGB-interleaved uint8 at arena + out_offsets[i]
// with row pitch 3 * widths[i].  Asynchronous on `stream` (nvJPEG copies the bitstreams itself).
MXB_API int MXBJpegDecodeBatch(void* h, const void* blob, const int64_t* offsets, const int64_t* lengths, int n, void* arena,
                               const int64_t* out_offsets, const int* widths, void* stream) {
  MXB_TRY
  auto* a = jpeg::api();
  auto* d = static_cast<jpeg::Decoder*>(h);
  if (n == 0) return 0;
  if (d->batch != n) {
    jpeg::check(a->BatchedInit(d->handle, d->state, n, d->threads, NVJPEG_OUTPUT_RGBI), "batched initialize");
    d->batch = n;
  }
  std::vector<const unsigned char*> data((size_t)n);
  std::vector<size_t> lens((size_t)n);
  std::vector<nvjpegImage_t> outs((size_t)n);
  for (int i = 0; i < n; ++i) {
    data[(size_t)i] = static_cast<const unsigned char*>(blob) + offsets[i];
    lens[(size_t)i] = (size_t)lengths[i];
    nvjpegImage_t im = {};
    im.channel[0] = static_cast<unsigned char*>(arena) + out_offsets[i];
    im.pitch[0] = 3 * (size_t)widths[i];
    outs[(size_t)i] = im;
  }
  const nvjpegStatus_t st = a->DecodeBatched(d->handle, d->state, data.data(), lens.data(), outs.data(), reinterpret_cast<cudaStream_t>(stream));
  if (st != NVJPEG_STATUS_SUCCESS) {
    d->batch = 0;     // a failed batch must be re-initialised
    jpeg::check(st, "decode batched");
  }
  MXB_CATCH
}

// windows: device array of n jpeg::Window; out: [n, OH, OW, 3] (layout 0) or [n, 3, OH, OW] (layout 1), dtype out_dt
MXB_API int MXBImageCropResize(const void* arena, const void* windows, int n, void* out, int out_dt, int layout, int OH, int OW,
                               const float* mean3, const float* std3, void* stream) {
  MXB_TRY
  if (n == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const dim3 grid((unsigned)((OH * OW + jpeg::kThreads - 1) / jpeg::kThreads), (unsigned)n);
  const float m0 = mean3[0], m1 = mean3[1], m2 = mean3[2], s0 = 1.f / std3[0], s1 = 1.f / std3[1], s2 = 1.f / std3[2];
  auto* w = static_cast<const jpeg::Window*>(windows);
  auto* ar = static_cast<const uint8_t*>(arena);
#define MXB_LAUNCH(T, L) jpeg::crop_resize_kernel<T, L><<<grid, jpeg::kThreads, 0, s>>>(ar, w, (T*)out, n, OH, OW, m0, m1, m2, s0, s1, s2)
  if (out_dt == kBF16 && layout == 0) MXB_LAUNCH(__nv_bfloat16, false);
  else if (out_dt == kBF16) MXB_LAUNCH(__nv_bfloat16, true);
  else if (out_dt == kF32 && layout == 0) MXB_LAUNCH(float, false);
  else if (out_dt == kF32) MXB_LAUNCH(float, true);
  else throw std::runtime_error("crop_resize: output must be bf16 or fp32");
#undef MXB_LAUNCH
  CUDA_CHECK(cudaGetLastError());
  MXB_CATCH
}
