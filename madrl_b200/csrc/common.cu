// Library-wide host helpers: error string, launch counter, device queries.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace madrl {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count(int device) {
  int n = 0;
  cudaError_t e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) {
    set_error("cudaDeviceGetAttribute(SM count): %s", cudaGetErrorString(e));
    return -1;
  }
  return n;
}

}  // namespace madrl

extern "C" const char* madrl_last_error(void) { return madrl::g_err; }
extern "C" int madrl_version(void) { return 100; }
extern "C" uint64_t madrl_launch_count(void) { return madrl::g_launches.load(); }

// ---- CUDA IPC helpers for the fused multi-GPU exchange (madrl_b200/dist.py PeerGather) -----------
// Buffers are cudaMalloc'ed here (not sub-allocated by a framework allocator) so that the IPC
// handle refers to exactly this buffer, and peers open it with THEIR device current, which is what
// makes the mapping usable from kernels running on the importing device.
extern "C" int madrl_ipc_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  MADRL_REQUIRE(ptr != nullptr && handle64 != nullptr && bytes > 0, "bad arguments");
  MADRL_CUDA_CHECK(cudaMalloc(ptr, bytes));
  MADRL_CUDA_CHECK(cudaMemset(*ptr, 0, bytes));
  cudaIpcMemHandle_t h;
  MADRL_CUDA_CHECK(cudaIpcGetMemHandle(&h, *ptr));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return MADRL_OK;
}

extern "C" int madrl_ipc_open(const unsigned char* handle64, void** ptr) {
  MADRL_REQUIRE(ptr != nullptr && handle64 != nullptr, "bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  MADRL_CUDA_CHECK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return MADRL_OK;
}

extern "C" int madrl_ipc_close(void* ptr) {
  if (ptr) MADRL_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
  return MADRL_OK;
}

extern "C" int madrl_ipc_free(void* ptr) {
  if (ptr) MADRL_CUDA_CHECK(cudaFree(ptr));
  return MADRL_OK;
}
