// Trajectory post-processing kernels on the rollout tensors: the
// consumers that sit right after the env hot path in the reference's sampler loop.  All three are
// pure HBM-bound streaming passes over [T][E][A][...] tensors, one thread per (env, agent[, dim])
// column walking the time axis, so every load / store is coalesced across the warp.
//
//   madrl_gae_f32            rllab/rllab/sampler/base.py:48-68 (deltas, GAE advantages, discounted
//                            returns via special.discount_cumsum, rllab/rllab/misc/special.py:107-111),
//                            paths segmented by `done`
//   madrl_frame_stack_f32    madrl_environments/__init__.py:143-196  ObservationBuffer
//   madrl_standardize_f32    madrl_environments/__init__.py:204-291  StandardizedEnv
#include "common.cuh"

namespace madrl {

// y[t] = x[t] + g * y[t+1] inside a path; a path ends where done[t] != 0 (bootstrap value 0, as
// base.py:57 appends 0) or at the end of the rollout (bootstrap = last_value or 0).
__global__ void gae_kernel(int T, int E, int A, const float* __restrict__ rew,
                           const float* __restrict__ val, const uint8_t* __restrict__ done,
                           const float* __restrict__ last_value, double discount, double gae_lambda,
                           float* __restrict__ adv, float* __restrict__ ret) {
  const size_t n = (size_t)E * A;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t e = i / A;
  double next_v = last_value ? (double)last_value[i] : 0.0;
  double run_adv = 0.0, run_ret = last_value ? (double)last_value[i] : 0.0;
  const double gl = discount * gae_lambda;
  for (int t = T - 1; t >= 0; --t) {
    const size_t k = (size_t)t * n + i;
    if (done[(size_t)t * E + e]) { next_v = 0.0; run_adv = 0.0; run_ret = 0.0; }
    const double r = (double)rew[k], v = (double)val[k];
    const double delta = r + discount * next_v - v;                // base.py:58-60
    run_adv = delta + gl * run_adv;                                // base.py:61-62
    run_ret = r + discount * run_ret;                              // base.py:63
    __stcs(adv + k, (float)run_adv);
    __stcs(ret + k, (float)run_ret);
    next_v = v;
  }
}

// ObservationBuffer: out[t][..][d][b] = the observation b steps before the newest one; on a step
// whose `done` flag is set the obs slot holds the reset observation and reset() fills every slot
// of the buffer with it (__init__.py:186-196).  `carry` [E][A][D][B] holds the buffer between calls.
template <int B>
__global__ void frame_stack_kernel(int T, size_t n /* E*A*D */, int AD, const float* __restrict__ obs,
                                   const uint8_t* __restrict__ done, int E, float* __restrict__ carry,
                                   float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t e = i / AD;
  float buf[B];
#pragma unroll
  for (int b = 0; b < B; ++b) buf[b] = carry[i * B + b];
  for (int t = 0; t < T; ++t) {
    const float x = obs[(size_t)t * n + i];
    if (done[(size_t)t * E + e]) {
#pragma unroll
      for (int b = 0; b < B; ++b) buf[b] = x;
    } else {
#pragma unroll
      for (int b = 0; b < B - 1; ++b) buf[b] = buf[b + 1];          // __init__.py:178-180
      buf[B - 1] = x;
    }
    float* o = out + ((size_t)t * n + i) * B;
#pragma unroll
    for (int b = 0; b < B; ++b) __stcs(o + b, buf[b]);
  }
#pragma unroll
  for (int b = 0; b < B; ++b) carry[i * B + b] = buf[b];
}

// StandardizedEnv running estimates (__init__.py:241-270), in place over the time axis.
//   mean <- (1-a) mean + a x ;  var <- (1-a) var + a (x - mean)^2 ;  out = (x - mean) / (sqrt(var) + eps)
// center = 1 for observations; rewards are only divided (and scaled), never centred.
__global__ void standardize_kernel(int T, size_t n, float* __restrict__ x, double* __restrict__ mean,
                                   double* __restrict__ var, double alpha, double eps, int center,
                                   double scale, int enable) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double m = mean[i], v = var[i];
  for (int t = 0; t < T; ++t) {
    const size_t k = (size_t)t * n + i;
    const double xv = (double)x[k];
    double o = xv;
    if (enable) {
      m = (1.0 - alpha) * m + alpha * xv;
      const double d = xv - m;
      v = (1.0 - alpha) * v + alpha * d * d;
      o = (center ? (xv - m) : xv) / (sqrt(v) + eps);
    }
    x[k] = (float)(scale * o);
  }
  mean[i] = m;
  var[i] = v;
}

// DiagnosticsWrapper episode statistics (madrl_environments/__init__.py:314-369): per env, walk the
// time axis accumulating the per-agent episode reward, the episode length and the discounted return
// of the agent-mean reward; an episode closes where done[t] is set or its length reaches
// max_traj_len (the wrapper then restarts its counters even though the env goes on).
// carry [E][A+3] doubles: episode reward per agent, length, discounted return, discount power.
__global__ void episode_stats_kernel(int T, int E, int A, const float* __restrict__ rew,
                                     const uint8_t* __restrict__ done, double discount, int max_traj_len,
                                     double* __restrict__ carry, float* __restrict__ ep_reward,
                                     float* __restrict__ ep_disc, int32_t* __restrict__ ep_len,
                                     uint8_t* __restrict__ ep_end) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  double* c = carry + (size_t)e * (A + 3);
  double len = c[A], disc = c[A + 1], pw = c[A + 2];
  if (len == 0.0) pw = 1.0;
  for (int t = 0; t < T; ++t) {
    const size_t k = ((size_t)t * E + e);
    double mean = 0.0;
    for (int a = 0; a < A; ++a) {
      const double r = (double)rew[k * A + a];
      c[a] += r;                                              // _episode_reward += rewardlist
      mean += r;
    }
    mean /= (double)A;
    disc += mean * pw;                                        // _discount_sum of the agent-mean reward
    pw *= discount;
    len += 1.0;
    const bool end = done[k] != 0 || len >= (double)max_traj_len;   // __init__.py:352
    ep_end[k] = end ? 1 : 0;
    ep_len[k] = end ? (int32_t)len : 0;
    ep_disc[k] = end ? (float)disc : 0.0f;
    for (int a = 0; a < A; ++a) ep_reward[k * A + a] = end ? (float)c[a] : 0.0f;
    if (end) {
      for (int a = 0; a < A; ++a) c[a] = 0.0;
      len = 0.0; disc = 0.0; pw = 1.0;
    }
  }
  c[A] = len; c[A + 1] = disc; c[A + 2] = pw;
}

}  // namespace madrl

using namespace madrl;

extern "C" int madrl_episode_stats_f32(int T, int E, int A, const float* rew_dev, const uint8_t* done_dev,
                                       double discount, int max_traj_len, double* carry_dev,
                                       float* ep_reward_dev, float* ep_disc_dev, int32_t* ep_len_dev,
                                       uint8_t* ep_end_dev, void* stream) {
  MADRL_REQUIRE(T >= 1 && E >= 1 && A >= 1 && max_traj_len >= 1, "bad sizes");
  MADRL_REQUIRE(rew_dev && done_dev && carry_dev && ep_reward_dev && ep_disc_dev && ep_len_dev && ep_end_dev,
                "NULL buffer");
  episode_stats_kernel<<<(unsigned)((E + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      T, E, A, rew_dev, done_dev, discount, max_traj_len, carry_dev, ep_reward_dev, ep_disc_dev, ep_len_dev,
      ep_end_dev);
  g_launches.fetch_add(1);
  MADRL_CUDA_CHECK(cudaGetLastError());
  return MADRL_OK;
}

extern "C" int madrl_gae_f32(int T, int E, int A, const float* rew_dev, const float* value_dev,
                             const uint8_t* done_dev, const float* last_value_dev, double discount,
                             double gae_lambda, float* adv_dev, float* ret_dev, void* stream) {
  MADRL_REQUIRE(T >= 1 && E >= 1 && A >= 1, "bad sizes");
  MADRL_REQUIRE(rew_dev && value_dev && done_dev && adv_dev && ret_dev, "NULL buffer");
  const size_t n = (size_t)E * A;
  gae_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      T, E, A, rew_dev, value_dev, done_dev, last_value_dev, discount, gae_lambda, adv_dev, ret_dev);
  g_launches.fetch_add(1);
  MADRL_CUDA_CHECK(cudaGetLastError());
  return MADRL_OK;
}

extern "C" int madrl_frame_stack_f32(int T, int E, int A, int D, int B, const float* obs_dev,
                                     const uint8_t* done_dev, float* carry_dev, float* out_dev,
                                     void* stream) {
  MADRL_REQUIRE(T >= 1 && E >= 1 && A >= 1 && D >= 1, "bad sizes");
  MADRL_REQUIRE(B >= 1 && B <= 8, "buffer_size must be in [1,8], got %d", B);
  MADRL_REQUIRE(obs_dev && done_dev && carry_dev && out_dev, "NULL buffer");
  const size_t n = (size_t)E * A * D;
  const unsigned grid = (unsigned)((n + 255) / 256);
  cudaStream_t s = (cudaStream_t)stream;
#define MADRL_FS(BB) case BB: frame_stack_kernel<BB><<<grid, 256, 0, s>>>(T, n, A * D, obs_dev, done_dev, E, carry_dev, out_dev); break
  switch (B) { MADRL_FS(1); MADRL_FS(2); MADRL_FS(3); MADRL_FS(4); MADRL_FS(5); MADRL_FS(6); MADRL_FS(7); MADRL_FS(8); }
#undef MADRL_FS
  g_launches.fetch_add(1);
  MADRL_CUDA_CHECK(cudaGetLastError());
  return MADRL_OK;
}

extern "C" int madrl_standardize_f32(int T, size_t n, float* x_dev, double* mean_dev, double* var_dev,
                                     double alpha, double eps, int center, double scale, int enable,
                                     void* stream) {
  MADRL_REQUIRE(T >= 1 && n >= 1, "bad sizes");
  MADRL_REQUIRE(x_dev && mean_dev && var_dev, "NULL buffer");
  standardize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      T, n, x_dev, mean_dev, var_dev, alpha, eps, center, scale, enable);
  g_launches.fetch_add(1);
  MADRL_CUDA_CHECK(cudaGetLastError());
  return MADRL_OK;
}
