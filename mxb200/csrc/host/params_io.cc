// Native .params writer (NDArray::Save list format, reference: src/ndarray/ndarray.cc:1574-1642,1774-1785).
//
//   file   := u64 0x112 | u64 0 | u64 n | n * blob | u64 m | m * (u64 len | bytes)
//   blobV2 := u32 0xF993fac9 | i32 stype(0 dense) | u32 ndim | ndim * i64 | i32 dev_type | i32 dev_id
//             | i32 type_flag | raw little-endian data
//
// Writing is done with one large buffered stream and an atomic rename so a crash never leaves a
// truncated checkpoint behind (the reference writes in place).  The python implementation in
// mxb200/ndarray/serialization.py is the format oracle; tests compare the two byte for byte.
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "../c_api.h"

namespace {
struct Writer {
  FILE* f;
  explicit Writer(FILE* fp) : f(fp) {}
  void raw(const void* p, size_t n) {
    if (n && fwrite(p, 1, n, f) != n) throw std::runtime_error("short write while saving .params");
  }
  template <typename T>
  void pod(T v) { raw(&v, sizeof(T)); }
};
}  // namespace

MXB_API int MXBParamsSave(const char* fname, int n, void** data, const int64_t* nbytes, const int* type_flags,
                          const int* ndims, const int64_t* shapes, const int* dev_types, const int* dev_ids,
                          int n_names, const char** names) {
  MXB_TRY
  MXB_CHECK(n_names == 0 || n_names == n, "names must be empty or match the number of arrays");
  std::string tmp = std::string(fname) + ".tmp." + std::to_string((long)getpid());
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) throw std::runtime_error(std::string("cannot open ") + tmp + " for writing");
  std::vector<char> iobuf(8 << 20);
  setvbuf(f, iobuf.data(), _IOFBF, iobuf.size());
  try {
    Writer w(f);
    w.pod<uint64_t>(0x112);
    w.pod<uint64_t>(0);
    w.pod<uint64_t>((uint64_t)n);
    const int64_t* sp = shapes;
    for (int i = 0; i < n; ++i) {
      w.pod<uint32_t>(0xF993fac9u);
      w.pod<int32_t>(0);
      w.pod<uint32_t>((uint32_t)ndims[i]);
      for (int d = 0; d < ndims[i]; ++d) w.pod<int64_t>(sp[d]);
      sp += ndims[i];
      if (ndims[i] == 0) continue;
      w.pod<int32_t>(dev_types[i]);
      w.pod<int32_t>(dev_ids[i]);
      w.pod<int32_t>(type_flags[i]);
      w.raw(data[i], (size_t)nbytes[i]);
    }
    w.pod<uint64_t>((uint64_t)n_names);
    for (int i = 0; i < n_names; ++i) {
      uint64_t len = strlen(names[i]);
      w.pod<uint64_t>(len);
      w.raw(names[i], len);
    }
  } catch (...) {
    fclose(f);
    unlink(tmp.c_str());
    throw;
  }
  if (fclose(f) != 0) {
    unlink(tmp.c_str());
    throw std::runtime_error("close failed while saving .params");
  }
  if (rename(tmp.c_str(), fname) != 0) {
    unlink(tmp.c_str());
    throw std::runtime_error(std::string("cannot rename checkpoint into ") + fname);
  }
  MXB_CATCH
}
