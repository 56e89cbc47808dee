// Stream-ordered primitives for the per-rollout multi-GPU exchange (madrl_b200/dist.py).
//
// Why they exist: the rollout kernel is a persistent wave that fills every SM.  A completion
// collective that runs as a KERNEL (an NCCL all-reduce of 4 bytes per rollout, round 1) needs SM
// slots from the next rollout's wave and spins on them while it waits for its peers; the driver
// measured the rollout kernel itself 2.8-3.8 % slower at N = 2..8 than at N = 1.  Everything here is
// executed by the copy engines and the stream front end instead -- no SM is involved:
//   madrl_copy_async        device->device cudaMemcpyAsync (CUDA-IPC peer mappings allowed: NVLink)
//   madrl_stream_write32    cuStreamWriteValue32 on LOCAL device memory
//   madrl_stream_wait_geq32 cuStreamWaitValue32(GEQ) on LOCAL device memory
// The flags live in the receiving GPU's own memory; remote GPUs set them with 4-byte peer copies
// that are stream-ordered behind the bulk copy they announce.
#include <cuda.h>

#include "common.cuh"

namespace {

typedef CUresult (*write32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*wait32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
write32_fn g_write32 = nullptr;
wait32_fn g_wait32 = nullptr;
bool g_resolved = false;

// The driver API is reached through the runtime (cudaGetDriverEntryPoint), so the library has no
// link-time dependency on libcuda.so and still loads on a machine without a driver (CPU tests).
int resolve() {
  if (g_resolved) return (g_write32 && g_wait32) ? MADRL_OK : MADRL_ECUDA;
  g_resolved = true;
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuStreamWriteValue32", &f, cudaEnableDefault, &q);
  if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) g_write32 = (write32_fn)f;
  f = nullptr;
  e = cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &q);
  if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) g_wait32 = (wait32_fn)f;
  if (!(g_write32 && g_wait32)) {
    madrl::set_error("stream memory operations are not available from this driver");
    return MADRL_ECUDA;
  }
  return MADRL_OK;
}

}  // namespace

extern "C" int madrl_stream_memops_available(void) { return resolve() == MADRL_OK ? 1 : 0; }

extern "C" int madrl_copy_async(void* dst_dev, const void* src_dev, size_t bytes, void* stream) {
  MADRL_REQUIRE(dst_dev != nullptr && src_dev != nullptr, "NULL pointer");
  if (bytes == 0) return MADRL_OK;
  MADRL_CUDA_CHECK(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return MADRL_OK;
}

extern "C" int madrl_stream_write32(void* stream, void* addr_dev, uint32_t value) {
  MADRL_REQUIRE(addr_dev != nullptr && ((uintptr_t)addr_dev & 3) == 0, "flag address must be 4-byte aligned");
  int rc = resolve();
  if (rc) return rc;
  CUresult r = g_write32((CUstream)stream, (CUdeviceptr)(uintptr_t)addr_dev, value, 0 /* default: with barrier */);
  if (r != CUDA_SUCCESS) { madrl::set_error("cuStreamWriteValue32 failed: %d", (int)r); return MADRL_ECUDA; }
  return MADRL_OK;
}

extern "C" int madrl_stream_wait_geq32(void* stream, void* addr_dev, uint32_t value) {
  MADRL_REQUIRE(addr_dev != nullptr && ((uintptr_t)addr_dev & 3) == 0, "flag address must be 4-byte aligned");
  int rc = resolve();
  if (rc) return rc;
  CUresult r = g_wait32((CUstream)stream, (CUdeviceptr)(uintptr_t)addr_dev, value, CU_STREAM_WAIT_VALUE_GEQ);
  if (r != CUDA_SUCCESS) { madrl::set_error("cuStreamWaitValue32 failed: %d", (int)r); return MADRL_ECUDA; }
  return MADRL_OK;
}
