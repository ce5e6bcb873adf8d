// Thread-local last-error storage for the host library (mirrors MXGetLastError, src/c_api/c_api_error.cc).
#include <string>

#include "../c_api.h"

namespace mxb {
static thread_local std::string g_last_error;
void set_last_error(const char* msg) { g_last_error = msg ? msg : ""; }
}  // namespace mxb

MXB_API const char* MXBGetLastError() { return mxb::g_last_error.c_str(); }
