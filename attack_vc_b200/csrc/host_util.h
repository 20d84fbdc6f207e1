// host_util.h -- error plumbing and device-memory arena shared by the host side of libavc_b200.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/avc_b200.h"

namespace avc {

struct Fail {
  int code;
  std::string msg;
};

[[noreturn]] inline void fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Fail{code, buf};
}

#define CK(expr)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess) fail(AVC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// ---- device memory owned by the handle / a plan ---------------------------------------------
struct Arena {
  // bump allocator over zero-initialised slabs: one cudaMalloc per 32 MiB instead of one per tensor
  static constexpr size_t kSlab = 32u << 20;
  std::vector<void*> slabs;
  char* cur = nullptr;
  size_t left = 0;
  size_t bytes = 0;
  Arena() = default;
  Arena(const Arena&) = delete;
  Arena& operator=(const Arena&) = delete;
  float* f(size_t n) {
    const size_t b = (n * sizeof(float) + 255) / 256 * 256;
    if (b > left) {
      const size_t sz = b > kSlab ? b : kSlab;
      void* p = nullptr;
      CK(cudaMalloc(&p, sz));
      CK(cudaMemset(p, 0, sz));
      slabs.push_back(p);
      bytes += sz;
      if (b > kSlab) return static_cast<float*>(p);   // dedicated slab, keep the current one
      cur = static_cast<char*>(p);
      left = sz;
    }
    float* r = reinterpret_cast<float*>(cur);
    cur += b;
    left -= b;
    return r;
  }
  template <class T>
  T* raw(size_t n) {
    return reinterpret_cast<T*>(f((n * sizeof(T) + 3) / 4));
  }
  float* upload(const std::vector<float>& v) {
    float* p = f(v.size());
    CK(cudaMemcpy(p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return p;
  }
  ~Arena() {
    for (void* p : slabs) cudaFree(p);
  }
};

}  // namespace avc
