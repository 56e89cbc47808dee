// TEST INFRASTRUCTURE -- not part of the product.
//
// A stand-in for <cuda_runtime.h> that lets g++ compile madrl_b200/csrc/{waterworld,pursuit,hostage,postproc,
// common}.cu unchanged, so that the `-m "not gpu"` test suite can execute the KERNEL SOURCE on the
// CPU and compare it with the oracle (tests/test_emulated_kernels.py).  Every CUDA thread of a block
// runs as a fiber (ucontext) on one OS thread; warp collectives (__shfl_sync, __ballot_sync,
// __reduce_or_sync, __syncwarp) and __syncthreads are rendezvous points between the fibers
// (tests/emu/emu_runtime.cpp).  "Device memory" is host memory; streams are synchronous.
//
// What this checks: the C++ semantics of the kernels and of the host-side launch code (indexing,
// masks, control flow, RNG draw order, state layout).  What it cannot check: what nvcc/ptxas make of
// them (FMA contraction, memory-model races between lanes that the lock-step fiber schedule hides).
// The GPU parity tests (`-m gpu`) remain the proof; nothing in madrl_b200/ loads this library.
#pragma once
#define MADRL_EMULATE 1

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <cmath>
#include <functional>
#include <type_traits>

// ---- qualifiers -------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#ifndef __shared__   // postproc.cu is compiled with -D__shared__=static (block-static arrays; blocks run one at a time)
#define __shared__
#endif
#define __align__(n) __attribute__((aligned(n)))

// ---- vector types -----------------------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct __attribute__((aligned(8))) float2 { float x, y; };
struct __attribute__((aligned(16))) double2 { double x, y; };
struct __attribute__((aligned(8))) int2 { int x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
static inline int2 make_int2(int x, int y) { int2 v; v.x = x; v.y = y; return v; }
static inline float2 make_float2(float x, float y) { float2 v; v.x = x; v.y = y; return v; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

// per-fiber built-ins: the scheduler rewrites them at every context switch
extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

// ---- runtime API (host memory, synchronous streams) ------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
typedef void* cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };

static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
typedef void* cudaEvent_t;
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = (cudaStream_t)(uintptr_t)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)(uintptr_t)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
// two "SMs" and one resident block each: a grid of two persistent blocks, so the env loop of the
// kernels (more envs than warps) is exercised.
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 2; return cudaSuccess; }
template <typename F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 1; return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
static inline cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaErrorNotSupported; }

// ---- fiber runtime (emu_runtime.cpp) --------------------------------------------------------------
namespace madrl_emu {
// Deposit `v` for this lane, wait until the 32 lanes of the warp have arrived, return the 32 slots.
const uint64_t* warp_exchange(uint64_t v);
void block_barrier();
void run_grid(unsigned grid, unsigned block, size_t smem, const std::function<void()>& body);
unsigned char* smem_anchor();   // origin of the 32-bit "shared addresses"
// Shared-memory hazard checker: every access through the helpers below is recorded per byte with
// the accessing lane and the number of __syncwarp/__syncthreads it has passed; two accesses to the
// same byte by different lanes, at least one a plain write (or a plain access vs an atomic), with no
// such barrier in between are a race on real hardware (lanes are not lock-step) even though the
// fiber schedule happens to order them.  Collectives other than __syncwarp/__syncthreads do not order
// memory.  Reports are counted and the first few printed; tests read the count.
enum SmemKind { SMEM_READ = 0, SMEM_WRITE = 1, SMEM_ATOMIC = 2 };
void smem_access(uint32_t addr, uint32_t bytes, SmemKind kind);
void note_sync();                // called by __syncwarp / __syncthreads of the current lane
extern "C" long madrl_emu_race_count();
extern "C" void madrl_emu_race_reset();
void check_smem(size_t bytes);

template <typename K, typename... A>
void launch(K kfn, unsigned grid, unsigned block, size_t smem, A... args) {
  run_grid(grid, block, smem, [=]() { kfn(args...); });
}
template <typename T> static inline uint64_t to_bits(T v) {
  static_assert(sizeof(T) <= 8, "collective payload too wide");
  uint64_t b = 0;
  memcpy(&b, &v, sizeof(T));
  return b;
}
template <typename T> static inline T from_bits(uint64_t b) {
  T v;
  memcpy(&v, &b, sizeof(T));
  return v;
}
}  // namespace madrl_emu

#define MADRL_LAUNCH(kfn, grid, block, smem, stream, ...) \
  madrl_emu::launch(kfn, (unsigned)(grid), (unsigned)(block), (size_t)(smem), __VA_ARGS__)

// ---- warp collectives (full masks only, as in the kernels) ------------------------------------------
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) {
  return madrl_emu::from_bits<T>(madrl_emu::warp_exchange(madrl_emu::to_bits(v))[src & 31]);
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) {
  return madrl_emu::from_bits<T>(madrl_emu::warp_exchange(madrl_emu::to_bits(v))[(threadIdx.x & 31) ^ m]);
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) {
  const int lane = threadIdx.x & 31, src = lane + d;
  const uint64_t* s = madrl_emu::warp_exchange(madrl_emu::to_bits(v));
  return madrl_emu::from_bits<T>(s[src < 32 ? src : lane]);
}
static inline unsigned __ballot_sync(unsigned, bool pred) {
  const uint64_t* s = madrl_emu::warp_exchange(pred ? 1u : 0u);
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) m |= (unsigned)(s[i] & 1u) << i;
  return m;
}
static inline unsigned __reduce_or_sync(unsigned, unsigned v) {
  const uint64_t* s = madrl_emu::warp_exchange(v);
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) m |= (unsigned)s[i];
  return m;
}
static inline unsigned __reduce_min_sync(unsigned, unsigned v) {
  const uint64_t* s = madrl_emu::warp_exchange(v);
  unsigned m = 0xffffffffu;
  for (int i = 0; i < 32; ++i) m = (unsigned)s[i] < m ? (unsigned)s[i] : m;
  return m;
}
static inline void __syncwarp(unsigned = 0xffffffffu) { madrl_emu::warp_exchange(0); madrl_emu::note_sync(); }
static inline void __syncthreads() { madrl_emu::block_barrier(); madrl_emu::note_sync(); }

// CUDA's global min / max overloads
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline float fminf_(float a, float b) { return a < b ? a : b; }

// ---- scalar intrinsics ---------------------------------------------------------------------------
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline float rsqrtf(float v) { return 1.0f / sqrtf(v); }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned sel) {   // PRMT, default mode
  const uint64_t src = ((uint64_t)y << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) r |= (unsigned)((src >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
}
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __float_as_uint(float v) { return (unsigned)madrl_emu::to_bits(v); }
static inline float __int_as_float(int v) { return madrl_emu::from_bits<float>((uint32_t)v); }
static inline double __longlong_as_double(long long v) { return madrl_emu::from_bits<double>((uint64_t)v); }
// round-to-nearest, never contracted into an FMA (volatile operands)
static inline double __dadd_rn(double a, double b) { volatile double x = a, y = b; return x + y; }
static inline double __dmul_rn(double a, double b) { volatile double x = a, y = b; return x * y; }
static inline double __ddiv_rn(double a, double b) { volatile double x = a, y = b; return x / y; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
template <typename T> static inline T __ldg(const T* p) { return *p; }

// ---- host versions of the PTX-level helpers of csrc/common.cuh -------------------------------------
#define MADRL_EMU_PTX_HELPERS 1
namespace madrl {
static inline char* emu_at(uint32_t a) { return reinterpret_cast<char*>(madrl_emu::smem_anchor()) + (int32_t)a; }
static inline uint32_t smem_addr(const void* p) {
  return (uint32_t)(reinterpret_cast<const char*>(p) - reinterpret_cast<const char*>(madrl_emu::smem_anchor()));
}
using madrl_emu::smem_access;
using madrl_emu::SMEM_READ; using madrl_emu::SMEM_WRITE; using madrl_emu::SMEM_ATOMIC;
static inline uint32_t lds_u32(uint32_t a) { smem_access(a, 4, SMEM_READ); uint32_t v; memcpy(&v, emu_at(a), 4); return v; }
static inline uint32_t lds_low_byte(uint32_t a) { smem_access(a, 1, SMEM_READ); return *reinterpret_cast<unsigned char*>(emu_at(a)); }
static inline float lds_f32(uint32_t a) { smem_access(a, 4, SMEM_READ); float v; memcpy(&v, emu_at(a), 4); return v; }
static inline uint32_t lds_u16(uint32_t a) { smem_access(a, 2, SMEM_READ); uint16_t v; memcpy(&v, emu_at(a), 2); return v; }
static inline void sts_u32(uint32_t a, uint32_t v) { smem_access(a, 4, SMEM_WRITE); memcpy(emu_at(a), &v, 4); }
static inline void sts_u16(uint32_t a, uint32_t v) { smem_access(a, 2, SMEM_WRITE); const uint16_t h = (uint16_t)v; memcpy(emu_at(a), &h, 2); }
static inline void reds_min_u32(uint32_t a, uint32_t v) {
  smem_access(a, 4, SMEM_ATOMIC);
  uint32_t w; memcpy(&w, emu_at(a), 4); w = v < w ? v : w; memcpy(emu_at(a), &w, 4);
}
static inline void reds_add_u32(uint32_t a, uint32_t v) {
  const uint32_t first = v ? (uint32_t)__builtin_ctz(v) / 8u : 0u;   // bytes below the lowest set bit of v never change
  smem_access(a + first, 4 - first, SMEM_ATOMIC);
  uint32_t w; memcpy(&w, emu_at(a), 4); w += v; memcpy(emu_at(a), &w, 4);
}
static inline void prefetch_l1(const void*) {}
static inline unsigned lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }

template <typename real> struct CandSlot;
template <> struct CandSlot<float> {
  static constexpr uint32_t kStride = 32;
  static inline void put(uint32_t a, float rx, float ry, float d2, float vx, float vy) {
    const float q[6] = {rx, ry, d2, d2, vx, vy};
    smem_access(a, sizeof q, SMEM_WRITE);
    memcpy(emu_at(a), q, sizeof q);
  }
  static inline void geom(uint32_t a, float& rx, float& ry, float& d2) {
    float q[3]; smem_access(a, sizeof q, SMEM_READ); memcpy(q, emu_at(a), sizeof q); rx = q[0]; ry = q[1]; d2 = q[2];
  }
  static inline void vel(uint32_t a, float& vx, float& vy) {
    float q[2]; smem_access(a + 16, sizeof q, SMEM_READ); memcpy(q, emu_at(a + 16), sizeof q); vx = q[0]; vy = q[1];
  }
};
template <> struct CandSlot<double> {
  static constexpr uint32_t kStride = 48;
  static inline void put(uint32_t a, double rx, double ry, double d2, double vx, double vy) {
    const double q[6] = {rx, ry, d2, 0.0, vx, vy};
    smem_access(a, sizeof q, SMEM_WRITE);
    memcpy(emu_at(a), q, sizeof q);
  }
  static inline void geom(uint32_t a, double& rx, double& ry, double& d2) {
    double q[3]; smem_access(a, sizeof q, SMEM_READ); memcpy(q, emu_at(a), sizeof q); rx = q[0]; ry = q[1]; d2 = q[2];
  }
  static inline void vel(uint32_t a, double& vx, double& vy) {
    double q[2]; smem_access(a + 32, sizeof q, SMEM_READ); memcpy(q, emu_at(a + 32), sizeof q); vx = q[0]; vy = q[1];
  }
};
}  // namespace madrl
