// Shared host/device helpers for the madrl_b200 C-ABI library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/madrl_b200.h"

namespace madrl {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
int sm_count(int device);

#define MADRL_CUDA_CHECK(expr)                                                          \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      madrl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                       __LINE__);                                                       \
      return MADRL_ECUDA;                                                               \
    }                                                                                   \
  } while (0)

#define MADRL_REQUIRE(cond, ...)      \
  do {                                \
    if (!(cond)) {                    \
      madrl::set_error(__VA_ARGS__);  \
      return MADRL_EINVAL;            \
    }                                 \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr unsigned FULL_MASK = 0xffffffffu;

template <typename real>
__device__ __forceinline__ real real_inf();
template <>
__device__ __forceinline__ float real_inf<float>() { return __int_as_float(0x7f800000); }
template <>
__device__ __forceinline__ double real_inf<double>() {
  return __longlong_as_double(0x7ff0000000000000ll);
}

// Streaming (evict-first) store for write-once trajectory tensors.
template <typename T>
__device__ __forceinline__ void store_stream(T* p, T v) { __stcs(p, v); }

template <typename real>
__device__ __forceinline__ real clip01(real x) {
  return x < (real)0 ? (real)0 : (x > (real)1 ? (real)1 : x);
}

}  // namespace madrl
