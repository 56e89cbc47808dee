// Library-wide host helpers: error string, launch counter, device queries.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace madrl {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<size_t> g_host_chunk_bytes{(size_t)32 << 20};

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
extern "C" int madrl_version(void) { return 200; }
// sizeof of every struct that crosses the ABI: a binding built against another header refuses to run
extern "C" void madrl_abi_sizes(int32_t* out6) {
  out6[0] = (int32_t)sizeof(madrl_ww_config); out6[1] = (int32_t)sizeof(madrl_ww_layout);
  out6[2] = (int32_t)sizeof(madrl_pursuit_config); out6[3] = (int32_t)sizeof(madrl_pursuit_layout);
  out6[4] = (int32_t)sizeof(madrl_hostage_config); out6[5] = (int32_t)sizeof(madrl_hostage_layout);
}
extern "C" uint64_t madrl_launch_count(void) { return madrl::g_launches.load(); }
extern "C" void madrl_set_host_chunk_bytes(size_t bytes) { madrl::g_host_chunk_bytes.store(bytes ? bytes : ((size_t)32 << 20)); }

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
