// The reference's hand-written policies as stand-alone action generators over a batch of observation rows
// (SURVEY.md 8f row 4).  The rollout kernels evaluate the same policies in-kernel on the features they have
// just computed (madrl_ww_rollout_heuristic / madrl_pursuit_rollout_heuristic: closed loop inside one
// launch); these entry points serve callers that step an env from the host, one observation batch at a time.
//
// Reference semantics: heuristics/waterworld.py:11-53 (hw:LINE), heuristics/pursuit.py:18-50 (hp:LINE).
#include <math.h>

#include "common.cuh"

namespace madrl {

// One warp per observation row.  obs [n][D] with the 7K feature layout of hw:12-22; act [n][2].
template <typename real>
__global__ void __launch_bounds__(128) ww_heuristic_kernel(size_t n, int K, int D, const real* __restrict__ obs,
                                                           real* __restrict__ act) {
  const int lane = threadIdx.x & 31;
  const size_t row = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;   // whole warps leave together
  const real* o = obs + row * (size_t)D;
  const real cE = o[7 * K] > (real)0 ? (real)1.5 : (real)1;        // hw:43
  const real cP = o[7 * K + 1] > (real)0 ? (real)1.5 : (real)1;    // hw:44
  const double step = (2.0 * M_PI - 0.0) / (double)K;              // hw:27 linspace(0, 2 pi, K + 1)[:-1]
  real wx = 0, wy = 0;
  for (int k = lane; k < K; k += 32) {
    const double a = (double)k * step;
    const real w = (real)0.5 * o[5 * K + k] + (cE * o[K + k] - o[k]) - cP * o[3 * K + k];   // hw:31-40,46
    wx += w * (real)cos(a);
    wy += w * (real)sin(a);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    wx += __shfl_xor_sync(FULL_MASK, wx, s);
    wy += __shfl_xor_sync(FULL_MASK, wy, s);
  }
  if (lane == 0) {
    // hw:47-51; scaled by the larger component first (the squares of a tiny sum would be denormal in float32)
    const real m = fmax(fabs(wx), fabs(wy));
    real ax = 0, ay = 0;
    if (m > (real)0) {
      const real ux = wx / m, uy = wy / m, nrm = sqrt(ux * ux + uy * uy);
      ax = ux / nrm; ay = uy / nrm;
    }
    act[2 * row] = ax;
    act[2 * row + 1] = ay;
  }
}

// One warp per observation row.  The evader channel of window cell w = wx * R + wy is obs[row][off + w * stride]
// (flatten: off = 2 R^2, stride 1; (R, R, 4) layout: off = 2, stride 4).  fallback [n]: the action taken when no
// evader is visible (hp:50 `action_space.sample()`, drawn by the caller).
__global__ void __launch_bounds__(128) pe_heuristic_kernel(size_t n, int R, int D, int off, int stride, int c2,
                                                           const float* __restrict__ obs, const int32_t* __restrict__ fallback,
                                                           const uint8_t* __restrict__ lut, int32_t* __restrict__ act) {
  const int lane = threadIdx.x & 31;
  const size_t row = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* o = obs + row * (size_t)D + off;
  uint32_t key = 0xffffffffu;
  for (int w = lane; w < R * R; w += 32) {
    if (o[(size_t)w * stride] > 0.0f) {
      const int ddx = 2 * (w / R) - c2, ddy = 2 * (w % R) - c2;     // hp:28-29 nearest, first in np.nonzero order
      key = min(key, ((uint32_t)(ddx * ddx + ddy * ddy) << 8) | (uint32_t)w);
    }
  }
  key = __reduce_min_sync(FULL_MASK, key);
  if (lane == 0) act[row] = key != 0xffffffffu ? (int32_t)lut[key & 0xffu] : fallback[row];
}

}  // namespace madrl

using namespace madrl;

void pe_policy_table(int R, int c2, uint8_t* lut);   // pursuit.cu

extern "C" int madrl_ww_heuristic_actions(int fp64, size_t n_rows, int n_sensors, int obs_dim, const void* obs_dev,
                                          void* actions_dev, void* stream) {
  MADRL_REQUIRE(obs_dev && actions_dev, "NULL buffer");
  MADRL_REQUIRE(n_sensors >= 1 && obs_dim >= 7 * n_sensors + 2, "obs_dim %d does not hold the 7K+2 layout for K = %d", obs_dim, n_sensors);
  if (n_rows == 0) return MADRL_OK;
  const unsigned grid = (unsigned)((n_rows + 3) / 4);
  if (fp64) MADRL_LAUNCH(ww_heuristic_kernel<double>, grid, 128, 0, (cudaStream_t)stream, n_rows, n_sensors, obs_dim,
                         (const double*)obs_dev, (double*)actions_dev);
  else MADRL_LAUNCH(ww_heuristic_kernel<float>, grid, 128, 0, (cudaStream_t)stream, n_rows, n_sensors, obs_dim,
                    (const float*)obs_dev, (float*)actions_dev);
  g_launches.fetch_add(1);
  MADRL_CUDA_CHECK(cudaGetLastError());
  return MADRL_OK;
}

extern "C" int madrl_pursuit_heuristic_actions(size_t n_rows, int obs_range, int flatten, int obs_dim, int floor_centre,
                                               const float* obs_dev, const int32_t* fallback_dev, uint8_t* lut_dev,
                                               int32_t* actions_dev, void* stream) {
  MADRL_REQUIRE(obs_dev && fallback_dev && lut_dev && actions_dev, "NULL buffer");
  const int R = obs_range, RR = R * R;
  MADRL_REQUIRE(R >= 1 && RR <= 128, "obs_range must be in [1,11]");
  MADRL_REQUIRE(obs_dim >= (flatten ? 3 * RR : 4 * RR), "obs_dim %d too small for obs_range %d", obs_dim, R);
  if (n_rows == 0) return MADRL_OK;
  uint8_t lut[128];
  const int c2 = floor_centre ? 2 * (R / 2) : R;
  pe_policy_table(R, c2, lut);
  MADRL_CUDA_CHECK(cudaMemcpyAsync(lut_dev, lut, (size_t)RR, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  const unsigned grid = (unsigned)((n_rows + 3) / 4);
  MADRL_LAUNCH(pe_heuristic_kernel, grid, 128, 0, (cudaStream_t)stream, n_rows, R, obs_dim, flatten ? 2 * RR : 2,
               flatten ? 1 : 4, c2, obs_dev, fallback_dev, (const uint8_t*)lut_dev, actions_dev);
  g_launches.fetch_add(1);
  MADRL_CUDA_CHECK(cudaGetLastError());
  return MADRL_OK;
}
