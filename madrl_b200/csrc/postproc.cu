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
// 1 / (sqrt(v) + eps) in float64 WITHOUT the DSQRT / DDIV sequences (each some two dozen FP64 instructions, which made
// these streaming passes FP64-bound at 29 % of the HBM roofline): float32 seeds (MUFU.RSQ / MUFU.RCP, ~1e-7) refined by
// one Newton step each in float64 (-> ~1e-14 relative; the result is rounded to float32 afterwards).  The running
// mean / variance recurrences themselves stay exact float64.  v == 0 (or denormal) takes the exact path.
__device__ __forceinline__ double inv_std(double v, double eps) {
  const float vf = (float)v;
  if (!(vf > 1e-30f) || !(vf < 1e30f)) return 1.0 / (sqrt(v) + eps);
  double y = (double)rsqrtf(vf);
  y = y * (1.5 - 0.5 * v * y * y);          // 1 / sqrt(v)
  const double den = v * y + eps;           // sqrt(v) + eps
  double r = (double)(1.0f / (float)den);
  r = r * (2.0 - den * r);
  return r;
}

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
      o = (center ? (xv - m) : xv) * inv_std(v, eps);
    }
    x[k] = (float)(scale * o);
  }
  mean[i] = m;
  var[i] = v;
}

// Observation standardiser for auto-reset rollouts with the terminal observations on the side.
__global__ void standardize_terminal_kernel(int T, int E, size_t per_env, float* __restrict__ x,
                                            float* __restrict__ term, const uint8_t* __restrict__ done,
                                            double* __restrict__ mean, double* __restrict__ var, double alpha,
                                            double eps) {
  const size_t n = (size_t)E * per_env;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t e = i / per_env;
  double m = mean[i], v = var[i];
  for (int t = 0; t < T; ++t) {
    const size_t k = (size_t)t * n + i;
    if (done[(size_t)t * E + e]) {                 // step(): the terminal observation comes first
      const double tv = (double)term[k];
      m = (1.0 - alpha) * m + alpha * tv;
      const double d = tv - m;
      v = (1.0 - alpha) * v + alpha * d * d;
      term[k] = (float)((tv - m) * inv_std(v, eps));
    }
    const double xv = (double)x[k];
    m = (1.0 - alpha) * m + alpha * xv;
    const double d = xv - m;
    v = (1.0 - alpha) * v + alpha * d * d;
    x[k] = (float)((xv - m) * inv_std(v, eps));
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


// ---------------------------------------------------------------------------------------------
// Whole-batch moments for advantage centring (rllab/rllab/algos/util.py:7-12, applied at
// rllab/rllab/sampler/base.py:82-86) and explained variance (rllab/rllab/misc/special.py:51-59).
// NumPy semantics: mean, then the population variance as the mean of squared deviations (two
// passes), float64 accumulation.  Deterministic: fixed grid, per-block partials combined in a fixed
// order by one warp (no atomics), so a given input always yields the same bits.
//
// Three series are derived per element from (a, b):  s0 = a,  s1 = b,  s2 = b - a  (b optional).
#ifndef MADRL_MOM_BLOCKS
#define MADRL_MOM_BLOCKS 592           // 4 CTAs x 148 SMs (the CPU emulator of tests/emu builds with fewer)
#endif
constexpr int kMomBlocks = MADRL_MOM_BLOCKS;
constexpr int kMomThreads = 256;
constexpr int kMomSeries = 3;
static_assert(kMomBlocks * 2 * kMomSeries <= MADRL_MOMENTS_WS, "workspace too small");

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_down_sync(0xffffffffu, v, o));
  return v;
}

// pass 0 (means == nullptr): partial sums and minima; pass 1: partial sums of squared deviations.
// part [kMomBlocks][2*kMomSeries].
__global__ void __launch_bounds__(kMomThreads) moments_partial_kernel(size_t n, const float* __restrict__ a,
                                                                      const float* __restrict__ b,
                                                                      const double* __restrict__ means,
                                                                      double* __restrict__ part) {
  double acc[kMomSeries] = {0.0, 0.0, 0.0};
  double mn[kMomSeries] = {INFINITY, INFINITY, INFINITY};
  double mu[kMomSeries] = {0.0, 0.0, 0.0};
  if (means) {
#pragma unroll
    for (int s = 0; s < kMomSeries; ++s) mu[s] = means[s];
  }
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double va = (double)__ldcs(a + i);
    const double vb = b ? (double)__ldcs(b + i) : 0.0;
    const double v[kMomSeries] = {va, vb, vb - va};
#pragma unroll
    for (int s = 0; s < kMomSeries; ++s) {
      if (means) {
        const double d = v[s] - mu[s];
        acc[s] += d * d;
      } else {
        acc[s] += v[s];
        mn[s] = fmin(mn[s], v[s]);
      }
    }
  }
  __shared__ double sh[kMomThreads / 32][2 * kMomSeries];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int s = 0; s < kMomSeries; ++s) {
    const double t = warp_sum(acc[s]);
    const double m = warp_min(mn[s]);
    if (lane == 0) { sh[w][2 * s] = t; sh[w][2 * s + 1] = m; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * kMomSeries) {
    double r = sh[0][threadIdx.x];
    for (int k = 1; k < kMomThreads / 32; ++k)
      r = (threadIdx.x & 1) ? fmin(r, sh[k][threadIdx.x]) : r + sh[k][threadIdx.x];
    part[(size_t)blockIdx.x * 2 * kMomSeries + threadIdx.x] = r;
  }
}

// One warp folds the block partials in a fixed order.  pass 0: stats[s] = mean_s, stats[6+s] = min_s;
// pass 1: stats[3+s] = var_s (population).
__global__ void moments_finalize_kernel(size_t n, const double* __restrict__ part, int pass,
                                        double* __restrict__ stats) {
  const int lane = threadIdx.x;
#pragma unroll
  for (int s = 0; s < kMomSeries; ++s) {
    double t = 0.0, m = INFINITY;
    for (int k = lane; k < kMomBlocks; k += 32) {
      t += part[(size_t)k * 2 * kMomSeries + 2 * s];
      m = fmin(m, part[(size_t)k * 2 * kMomSeries + 2 * s + 1]);
    }
    t = warp_sum(t);
    m = warp_min(m);
    if (lane == 0) {
      if (pass == 0) { stats[s] = t / (double)n; stats[6 + s] = m; }
      else stats[3 + s] = t / (double)n;
    }
  }
}

// center_advantages: (x - mean) / (std + 1e-8); shift_advantages_to_positive: (x - min) + 1e-8 on the
// (possibly centred) values, util.py:7-12.
__global__ void center_apply_kernel(size_t n, float* __restrict__ x, const double* __restrict__ stats,
                                    int center, int positive) {
  const double mean = stats[0], sd = sqrt(stats[3]) + 1e-8, mn = stats[6];
  const double cmin = center ? (mn - mean) / sd : mn;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double v = (double)x[i];
    if (center) v = (v - mean) / sd;
    if (positive) v = (v - cmin) + 1e-8;
    __stcs(x + i, (float)v);
  }
}


// ---- path packing: time-major rollout tensors -> rllab's per-(env, episode, agent) paths --------
// (rllab/rllab/sampler/ma_sampler.py:52-100 dec_rollout returns one path dict per agent with
// `observations / actions / rewards / env_infos` arrays; VecEnvExecutor cuts episodes at `done`,
// rllab/sandbox/rocky/tf/envs/vec_env_executor.py:16-28.)  Every (t, env, agent) row of a rollout
// belongs to exactly one path, so packing is a permutation of rows: paths are ordered (env, episode,
// agent) like the host `to_paths`, the rows of a path are consecutive, and row (t, e, a) lands at
//     e*T*A + s*A + a*L + (t - s)        s, L = start and length of the episode containing t.
// paths_plan_kernel: one thread per env walks `done` once (T bytes) and writes s / L per (t, env),
// the per-episode records and the episode count; paths_pack_kernel: one warp per source row, coalesced
// reads and writes of D 4-byte words -- a pure HBM-bound permutation (2 x the tensor's bytes).
__global__ void paths_plan_kernel(int T, int E, int A, const uint8_t* __restrict__ done,
                                  int32_t* __restrict__ seg_start, int32_t* __restrict__ seg_len,
                                  int32_t* __restrict__ n_episodes, int32_t* __restrict__ ep_rec) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int s = 0, n = 0;
  for (int t = 0; t < T; ++t) {
    const bool end = done[(size_t)t * E + e] != 0;
    if (end || t == T - 1) {
      const int L = t + 1 - s;
      for (int u = s; u <= t; ++u) { seg_start[(size_t)u * E + e] = s; seg_len[(size_t)u * E + e] = L; }
      // episode record n of this env (at most T episodes): start step, length, terminated
      int32_t* r = ep_rec + 3 * ((size_t)e * T + n);
      r[0] = s; r[1] = L; r[2] = end ? 1 : 0;
      ++n;
      s = t + 1;
    }
  }
  n_episodes[e] = n;
}

__global__ void paths_pack_kernel(int T, int E, int A, int D, const uint32_t* __restrict__ src,
                                  const uint32_t* __restrict__ first, const int32_t* __restrict__ seg_start,
                                  const int32_t* __restrict__ seg_len, uint32_t* __restrict__ dst) {
  const size_t row = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // source row (t, e, a)
  const int lane = threadIdx.x & 31;
  const size_t rows = (size_t)T * E * A;
  if (row >= rows) return;
  const int a = (int)(row % A);
  const size_t te = row / A;
  const int e = (int)(te % E), t = (int)(te / E);
  const int s = seg_start[te], L = seg_len[te];
  // `first` != NULL: observations are shifted by one step -- the row of time t is the observation the
  // action of time t was taken in: first[e][a] for t = 0, src[t-1] afterwards (on a done step the
  // rollout's obs slot already holds the reset observation, i.e. the next episode's first one)
  const uint32_t* in = first ? (t == 0 ? first + ((size_t)e * A + a) * D : src + (row - (size_t)E * A) * D)
                             : src + row * D;
  uint32_t* out = dst + ((size_t)e * T * A + (size_t)s * A + (size_t)a * L + (t - s)) * D;
  for (int d = lane; d < D; d += 32) __stcs(out + d, __ldcs(in + d));
}

// Narrow rows (rewards, actions, infos: D <= 4 words): one THREAD per source row instead of one warp.
__global__ void paths_pack_narrow_kernel(int T, int E, int A, int D, const uint32_t* __restrict__ src,
                                         const uint32_t* __restrict__ first, const int32_t* __restrict__ seg_start,
                                         const int32_t* __restrict__ seg_len, uint32_t* __restrict__ dst) {
  const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // source row (t, e, a)
  const size_t rows = (size_t)T * E * A;
  if (row >= rows) return;
  const int a = (int)(row % A);
  const size_t te = row / A;
  const int e = (int)(te % E), t = (int)(te / E);
  const int s = seg_start[te], L = seg_len[te];
  const uint32_t* in = first ? (t == 0 ? first + ((size_t)e * A + a) * D : src + (row - (size_t)E * A) * D)
                             : src + row * D;
  uint32_t* out = dst + ((size_t)e * T * A + (size_t)s * A + (size_t)a * L + (t - s)) * D;
  for (int d = 0; d < D; ++d) out[d] = in[d];
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
  MADRL_LAUNCH(episode_stats_kernel, (unsigned)((E + 127) / 128), 128, 0, (cudaStream_t)stream, 
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
  MADRL_LAUNCH(gae_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, 
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
#define MADRL_FS(BB) case BB: MADRL_LAUNCH(frame_stack_kernel<BB>, grid, 256, 0, s, T, n, A * D, obs_dev, done_dev, E, carry_dev, out_dev); break
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
  MADRL_LAUNCH(standardize_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, 
      T, n, x_dev, mean_dev, var_dev, alpha, eps, center, scale, enable);
  g_launches.fetch_add(1);
  MADRL_CUDA_CHECK(cudaGetLastError());
  return MADRL_OK;
}

static int run_moments(size_t n, const float* a, const float* b, double* stats, double* ws, cudaStream_t s) {
  MADRL_LAUNCH(moments_partial_kernel, kMomBlocks, kMomThreads, 0, s, n, a, b, nullptr, ws);
  MADRL_LAUNCH(moments_finalize_kernel, 1, 32, 0, s, n, ws, 0, stats);
  MADRL_LAUNCH(moments_partial_kernel, kMomBlocks, kMomThreads, 0, s, n, a, b, stats, ws);
  MADRL_LAUNCH(moments_finalize_kernel, 1, 32, 0, s, n, ws, 1, stats);
  g_launches.fetch_add(4);
  MADRL_CUDA_CHECK(cudaGetLastError());
  return MADRL_OK;
}

extern "C" int madrl_center_advantages_f32(size_t n, float* adv_dev, int center, int positive,
                                           double* stats_dev, double* workspace_dev, void* stream) {
  MADRL_REQUIRE(n >= 1, "bad sizes");
  MADRL_REQUIRE(adv_dev && stats_dev && workspace_dev, "NULL buffer");
  cudaStream_t s = (cudaStream_t)stream;
  const int rc = run_moments(n, adv_dev, nullptr, stats_dev, workspace_dev, s);
  if (rc) return rc;
  if (center || positive) {
    MADRL_LAUNCH(center_apply_kernel, kMomBlocks, kMomThreads, 0, s, n, adv_dev, stats_dev, center, positive);
    g_launches.fetch_add(1);
    MADRL_CUDA_CHECK(cudaGetLastError());
  }
  return MADRL_OK;
}

extern "C" int madrl_moments_f32(size_t n, const float* a_dev, const float* b_dev, double* stats_dev,
                                 double* workspace_dev, void* stream) {
  MADRL_REQUIRE(n >= 1, "bad sizes");
  MADRL_REQUIRE(a_dev && stats_dev && workspace_dev, "NULL buffer");
  return run_moments(n, a_dev, b_dev, stats_dev, workspace_dev, (cudaStream_t)stream);
}

extern "C" int madrl_paths_plan(int T, int E, int A, const uint8_t* done_dev, int32_t* seg_start_dev,
                                int32_t* seg_len_dev, int32_t* n_episodes_dev, int32_t* ep_rec_dev, void* stream) {
  MADRL_REQUIRE(T >= 1 && E >= 1 && A >= 1, "bad shape");
  MADRL_REQUIRE((size_t)T * E * A < ((size_t)1 << 31), "T*E*A must be < 2^31 (row offsets are int32)");
  MADRL_REQUIRE(done_dev && seg_start_dev && seg_len_dev && n_episodes_dev && ep_rec_dev, "NULL pointer");
  MADRL_LAUNCH(paths_plan_kernel, (E + 127) / 128, 128, 0, (cudaStream_t)stream, T, E, A, done_dev, seg_start_dev,
               seg_len_dev, n_episodes_dev, ep_rec_dev);
  g_launches.fetch_add(1);
  MADRL_CUDA_CHECK(cudaGetLastError());
  return MADRL_OK;
}

extern "C" int madrl_paths_pack_u32(int T, int E, int A, int D, const void* src_dev, const void* first_dev,
                                    const int32_t* seg_start_dev, const int32_t* seg_len_dev, void* dst_dev,
                                    void* stream) {
  MADRL_REQUIRE(T >= 1 && E >= 1 && A >= 1 && D >= 1, "bad shape");
  MADRL_REQUIRE(src_dev && seg_start_dev && seg_len_dev && dst_dev, "NULL pointer");
  const size_t rows = (size_t)T * E * A;
  const size_t blocks = D <= 4 ? (rows + 255) / 256 : (rows * 32 + 255) / 256;
  MADRL_REQUIRE(blocks < ((size_t)1 << 31), "too many rows for one launch");
  if (D <= 4)
    MADRL_LAUNCH(paths_pack_narrow_kernel, (unsigned)blocks, 256, 0, (cudaStream_t)stream, T, E, A, D,
                 (const uint32_t*)src_dev, (const uint32_t*)first_dev, seg_start_dev, seg_len_dev, (uint32_t*)dst_dev);
  else
    MADRL_LAUNCH(paths_pack_kernel, (unsigned)blocks, 256, 0, (cudaStream_t)stream, T, E, A, D, (const uint32_t*)src_dev,
                 (const uint32_t*)first_dev, seg_start_dev, seg_len_dev, (uint32_t*)dst_dev);
  g_launches.fetch_add(1);
  MADRL_CUDA_CHECK(cudaGetLastError());
  return MADRL_OK;
}

extern "C" int madrl_standardize_obs_terminal_f32(int T, int E, size_t per_env, float* x_dev, float* term_dev,
                                                  const uint8_t* done_dev, double* mean_dev, double* var_dev,
                                                  double alpha, double eps, void* stream) {
  MADRL_REQUIRE(T >= 1 && E >= 1 && per_env >= 1, "bad sizes");
  MADRL_REQUIRE(x_dev && term_dev && done_dev && mean_dev && var_dev, "NULL buffer");
  const size_t n = (size_t)E * per_env;
  MADRL_LAUNCH(standardize_terminal_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, T, E, per_env,
               x_dev, term_dev, done_dev, mean_dev, var_dev, alpha, eps);
  g_launches.fetch_add(1);
  MADRL_CUDA_CHECK(cudaGetLastError());
  return MADRL_OK;
}
