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
