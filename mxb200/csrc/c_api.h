// Flat C ABI shared by libmxb200_host.so and libmxb200_cuda.so.
//
// Parity: the reference exposes its runtime through include/mxnet/c_api.h and reports
// failures with MXGetLastError(); python binds it with ctypes (python/mxnet/base.py).
// We keep that layering (it also keeps every translation unit free of torch headers).
// Every entry point returns 0 on success, non-zero on failure; the message is then
// available from MXBGetLastError() (thread-local).
#pragma once
#include <stdint.h>
#include <stdexcept>
#include <string>

#define MXB_API extern "C" __attribute__((visibility("default")))

namespace mxb {
void set_last_error(const char* msg);
}

#define MXB_TRY try {
#define MXB_CATCH                                   \
  }                                                 \
  catch (const std::exception& e) {                 \
    mxb::set_last_error(e.what());                  \
    return -1;                                      \
  }                                                 \
  catch (...) {                                     \
    mxb::set_last_error("unknown native exception"); \
    return -1;                                      \
  }                                                 \
  return 0;

#define MXB_CHECK(cond, msg)                                                        \
  do {                                                                              \
    if (!(cond)) throw std::runtime_error(std::string("Check failed: " #cond " ") + (msg)); \
  } while (0)

MXB_API const char* MXBGetLastError();
