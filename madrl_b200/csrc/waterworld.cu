// MAWaterWorld batched engine: persistent multi-step rollout kernel for sm_100a.
//
// Reference semantics: madrl_environments/pursuit/waterworld.py (cited as ww:LINE).
//
// Design:
//   * HBM state is struct-of-arrays, env index minor: pos/vel [object][E], so a block's
//     contiguous env range loads/stores short coalesced runs; trajectory tensors are
//     [T][E][Np][D] (agent rows contiguous -- the layout the reference callers consume).
//   * A block owns a contiguous range of envs for the WHOLE T-step rollout.  Their state is
//     staged into shared memory once and never touches HBM again until the rollout ends;
//     per step the only HBM traffic is actions in and obs / reward / done / info out.
//   * Per step, two block phases:
//       "env phase"  one warp per env, lanes over objects: integrate + walls + obstacle
//                    rebound (ww:229-270) for step t, fused with the tail of step t-1
//                    (catch bookkeeping, respawn, rewards, evader drift, ww:358-409,433-436);
//       "sense phase" one warp per (env, pursuer), lanes over objects then over sensors:
//                    pairwise collisions via __ballot_sync (ww:278-293), exact conservative
//                    range cull via ballot, nearest-object-per-sensor over the surviving
//                    candidates (ww:64-72,312-353), coalesced feature-major obs row stores
//                    (ww:388-428).
//   * fp32 production instantiation and fp64 verification instantiation of the same template.
#include <math.h>
#include <new>

#include "common.cuh"
#include "philox.cuh"

namespace madrl {

template <typename real>
struct WWParams {
  int E, env_id_base, Np, Ne, Npo, K, n_coop, D, Nall, CW;
  int reward_global, addid, speed_features, random_obstacle, timestep_limit, max_path_length;
  int T, mode, auto_reset;  // mode 0 = rollout, 1 = reset
  int max_loc, env_stride;  // envs per block (max), bytes per env slot in smem
  int off_py, off_vx, off_vy, off_rew, off_coll, off_meta, off_ctr;  // byte offsets in a slot
  real r_p2, range, cull2, rsum_e, rsum_po;          // sensing / collision thresholds
  real thr_obst_p, thr_obst_e, thr_obst_po;          // r_class + obstacle_radius (ww:251,259,267)
  real thr_resp_p, thr_resp_e, thr_resp_po;          // 2 r_class + obstacle_radius (ww:140)
  real obst_x, obst_y, ev_speed, poison_speed, action_scale;
  real poison_reward, food_reward, encounter_reward, control_penalty;
  uint64_t seed;
  // state (SoA, env minor)
  real *pos_x, *pos_y, *vel_x, *vel_y, *obst_px, *obst_py;
  int32_t *timestep, *path_len;
  uint64_t* ctr;
  const real* sensors;
  // trajectory tensors
  const real* actions;
  real* obs;
  real* rew;
  uint8_t* done;
  int32_t* info;
  const uint8_t* mask;
};

template <typename real> struct Vec2;
template <> struct Vec2<float> { typedef float2 type; };
template <> struct Vec2<double> { typedef double2 type; };

template <typename real>
struct EnvSlot {
  real *px, *py, *vx, *vy, *rew;
  uint32_t* coll;  // [Np][CW] bit o set <=> pursuer row collides with evader/poison object o
  int32_t* meta;   // 0: timestep, 1: path_len, 2: needs_reset
  uint64_t* ctr;
};

template <typename real>
__device__ __forceinline__ EnvSlot<real> env_slot(const WWParams<real>& p, char* base, int slot) {
  char* b = base + (size_t)slot * p.env_stride;
  EnvSlot<real> s;
  s.px = reinterpret_cast<real*>(b);
  s.py = reinterpret_cast<real*>(b + p.off_py);
  s.vx = reinterpret_cast<real*>(b + p.off_vx);
  s.vy = reinterpret_cast<real*>(b + p.off_vy);
  s.rew = reinterpret_cast<real*>(b + p.off_rew);
  s.coll = reinterpret_cast<uint32_t*>(b + p.off_coll);
  s.meta = reinterpret_cast<int32_t*>(b + p.off_meta);
  s.ctr = reinterpret_cast<uint64_t*>(b + p.off_ctr);
  return s;
}

template <typename real>
__device__ __forceinline__ real warp_sum(real v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}

// ww:139-142  rejection loop around the obstacle (lane-serial, rare).
template <typename real>
__device__ __forceinline__ void spawn_point(SeqStream& rs, real obx, real oby, real thr, real& x,
                                            real& y) {
  x = rs.next_unit<real>();
  y = rs.next_unit<real>();
  while (true) {
    const real dx = x - obx, dy = y - oby;
    if (!(sqrt(dx * dx + dy * dy) <= thr)) break;
    x = rs.next_unit<real>();
    y = rs.next_unit<real>();
  }
}

// ww:144-170  reset draws, in the reference's order (obstacle, pursuers, evaders, poisons).
template <typename real>
__device__ void env_reset_draws(const WWParams<real>& p, const EnvSlot<real>& s, int e, int lane) {
  if (lane == 0) {
    SeqStream rs;
    rs.init(p.seed, (uint32_t)(p.env_id_base + e), 0u, *s.ctr);
    real obx = p.obst_x, oby = p.obst_y;
    if (p.random_obstacle) {
      obx = rs.next_unit<real>();
      oby = rs.next_unit<real>();
    }
    s.px[p.Nall] = obx;
    s.py[p.Nall] = oby;
    for (int o = 0; o < p.Nall; ++o) {
      const bool isP = o < p.Np, isE = !isP && o < p.Np + p.Ne;
      real x, y;
      spawn_point<real>(rs, obx, oby, isP ? p.thr_resp_p : (isE ? p.thr_resp_e : p.thr_resp_po), x, y);
      real vx = 0, vy = 0;
      if (!isP) {  // ww:164,170 -- poisons also use ev_speed at reset
        vx = (rs.next_unit<real>() - (real)0.5) * p.ev_speed;
        vy = (rs.next_unit<real>() - (real)0.5) * p.ev_speed;
      }
      s.px[o] = x; s.py[o] = y; s.vx[o] = vx; s.vy[o] = vy;
    }
    *s.ctr = rs.counter;
    s.meta[0] = 0;
    s.meta[1] = 0;
  }
  __syncwarp();
}

// ww:221-270  integrate pursuers, control penalty, walls, obstacle rebound (all objects).
template <typename real>
__device__ void env_pre(const WWParams<real>& p, const EnvSlot<real>& s, int e, int t,
                        bool zero_action, int lane) {
  const real obx = s.px[p.Nall], oby = s.py[p.Nall];
  real sq_tot = 0;
  for (int base = 0; base < p.Nall; base += 32) {
    const int o = base + lane;
    real sq = 0;
    if (o < p.Nall) {
      real x = s.px[o], y = s.py[o], vx = s.vx[o], vy = s.vy[o];
      real thr, k;
      if (o < p.Np) {
        real ax = 0, ay = 0;
        if (!zero_action) {
          typedef typename Vec2<real>::type V2;
          const V2 a = reinterpret_cast<const V2*>(p.actions)[((size_t)t * p.E + e) * p.Np + o];
          ax = a.x * p.action_scale;
          ay = a.y * p.action_scale;
        }
        vx += ax; vy += ay;
        x += vx;  y += vy;
        sq = ax * ax + ay * ay;
        const real cx = clip01(x), cy = clip01(y);
        if (x != cx) vx = 0;
        if (y != cy) vy = 0;
        x = cx; y = cy;
        thr = p.thr_obst_p; k = (real)-0.5;
        if (!p.reward_global) s.rew[o] = p.control_penalty * sq;
      } else if (o < p.Np + p.Ne) {
        thr = p.thr_obst_e; k = (real)-0.5;
      } else {
        thr = p.thr_obst_po; k = (real)-1;
      }
      const real dx = x - obx, dy = y - oby;
      if (sqrt(dx * dx + dy * dy) <= thr) { vx = k * vx; vy = k * vy; }
      s.px[o] = x; s.py[o] = y; s.vx[o] = vx; s.vy[o] = vy;
    }
    if (p.reward_global && base < p.Np) sq_tot += warp_sum(sq);
  }
  if (p.reward_global) {
    for (int o = lane; o < p.Np; o += 32) s.rew[o] = p.control_penalty * sq_tot;
  }
}

// ww:285,293,358-385,397-409,433-436  catches, respawn, rewards, drift, bookkeeping.
template <typename real>
__device__ void env_post(const WWParams<real>& p, const EnvSlot<real>& s, int e, int t,
                         bool discard, int lane) {
  SeqStream rs;
  rs.init(p.seed, (uint32_t)(p.env_id_base + e), 0u, *s.ctr);
  const real obx = s.px[p.Nall], oby = s.py[p.Nall];
  unsigned whoE = 0, whoP = 0, whoEnc = 0;
  int nE = 0, nP = 0, nEnc = 0;
  for (int base = p.Np; base < p.Nall; base += 32) {
    const int o = base + lane;
    const bool v = o < p.Nall;
    unsigned col = 0;
    if (v) {
      for (int q = 0; q < p.Np; ++q) col |= ((s.coll[q * p.CW + (o >> 5)] >> (o & 31)) & 1u) << q;
    }
    const int cnt = __popc(col);
    const bool isE = o < p.Np + p.Ne;
    const bool caught = v && (isE ? cnt >= p.n_coop : cnt >= 1);   // ww:285 / ww:293
    const bool enc = v && isE && cnt >= 1;                          // ww:376
    const unsigned cm = __ballot_sync(FULL_MASK, caught);
    const unsigned cmE = __ballot_sync(FULL_MASK, caught && isE);
    nE += __popc(cmE);
    nP += __popc(cm & ~cmE);
    nEnc += __popc(__ballot_sync(FULL_MASK, enc));
    if (caught) { if (isE) whoE |= col; else whoP |= col; }
    if (enc) whoEnc |= col;
    if (cm != 0u) {
      if (lane == 0) {  // ww:358-374, ascending object index = evaders first, then poisons
        unsigned m = cm;
        while (m) {
          const int j = __ffs(m) - 1;
          m &= m - 1;
          const int oj = base + j;
          const bool jE = oj < p.Np + p.Ne;
          real x, y;
          spawn_point<real>(rs, obx, oby, jE ? p.thr_resp_e : p.thr_resp_po, x, y);
          const real sp = jE ? p.ev_speed : p.poison_speed;
          s.px[oj] = x; s.py[oj] = y;
          s.vx[oj] = (rs.next_unit<real>() - (real)0.5) * sp;
          s.vy[oj] = (rs.next_unit<real>() - (real)0.5) * sp;
        }
      }
      __syncwarp();
    }
  }
  whoE = __reduce_or_sync(FULL_MASK, whoE);
  whoP = __reduce_or_sync(FULL_MASK, whoP);
  whoEnc = __reduce_or_sync(FULL_MASK, whoEnc);
  // rewards ww:376-385
  if (!discard) {
    for (int q = lane; q < p.Np; q += 32) {
      real r = s.rew[q];
      if (p.reward_global) {
        r += ((real)nE * p.food_reward + (real)nP * p.poison_reward) + (real)nEnc * p.encounter_reward;
      } else {
        if ((whoE >> q) & 1u) r += p.food_reward;
        if ((whoP >> q) & 1u) r += p.poison_reward;
        if ((whoEnc >> q) & 1u) r += p.encounter_reward;
      }
      store_stream(p.rew + ((size_t)t * p.E + e) * p.Np + q, r);
    }
  }
  // evaders / poison drift, ww:397-409
  for (int o = p.Np + lane; o < p.Nall; o += 32) {
    real x = s.px[o] + s.vx[o], y = s.py[o] + s.vy[o];
    s.px[o] = x; s.py[o] = y;
    const bool ox = (x < (real)0) || (x > (real)1), oy = (y < (real)0) || (y > (real)1);
    if (ox && oy) { s.vx[o] = -s.vx[o]; s.vy[o] = -s.vy[o]; }
  }
  if (lane == 0) {
    *s.ctr = rs.counter;
    const int tt = s.meta[0] + 1, ts = s.meta[1] + 1;  // ww:433
    s.meta[0] = tt;
    s.meta[1] = ts;
    const bool done = (tt >= p.timestep_limit) || (p.max_path_length > 0 && ts >= p.max_path_length);
    if (!discard) {
      p.done[(size_t)t * p.E + e] = done ? 1 : 0;
      p.info[((size_t)t * p.E + e) * 2 + 0] = nE;
      p.info[((size_t)t * p.E + e) * 2 + 1] = nP;
      s.meta[2] = (done && p.auto_reset) ? 1 : 0;
    } else {
      s.meta[2] = 0;
    }
  }
  __syncwarp();
}

// ww:64-72,278-353,388-428  one (env, pursuer): collisions, sensing, obs row.
template <typename real>
__device__ void sense_item(const WWParams<real>& p, const EnvSlot<real>& s, const real* sens,
                           int pi, int lane, real* __restrict__ obs_row) {
  const real INF = real_inf<real>();
  const real mx = s.px[pi], my = s.py[pi], mvx = s.vx[pi], mvy = s.vy[pi];
  const int Nobj = p.Nall + 1;  // obstacle is object index Nall
  const int eLo = p.Np, eHi = p.Np + p.Ne;
  bool anyE = false, anyP = false;
  for (int ks = 0; ks < p.K; ks += 32) {
    const int k = ks + lane;
    const bool kval = k < p.K;
    const real sx = kval ? sens[k] : (real)0, sy = kval ? sens[p.K + k] : (real)0;
    real bU = INF, bE = INF, bP = INF, bO = INF;
    int iU = 0, iE = 0, iP = 0;
    for (int base = 0; base < Nobj; base += 32) {
      const int o = base + lane;
      const bool valid = o < Nobj;
      real d2 = INF;
      if (valid) {
        const real rx = s.px[o] - mx, ry = s.py[o] - my;
        d2 = rx * rx + ry * ry;
      }
      unsigned cm = __ballot_sync(FULL_MASK, valid && (d2 <= p.cull2) && (o != pi));
      if (ks == 0) {  // pairwise collisions, ww:278-293
        const bool isE = o >= eLo && o < eHi, isP = o >= eHi && o < p.Nall;
        const real dist = sqrt(d2);
        const bool c = (isE && dist <= p.rsum_e) || (isP && dist <= p.rsum_po);
        const unsigned m = __ballot_sync(FULL_MASK, c);
        anyE |= __ballot_sync(FULL_MASK, c && isE) != 0u;
        anyP |= __ballot_sync(FULL_MASK, c && isP) != 0u;
        if (lane == 0) s.coll[pi * p.CW + (base >> 5)] = m;
      }
      while (cm) {  // warp-uniform loop over in-range candidates, ascending index
        const int j = __ffs(cm) - 1;
        cm &= cm - 1;
        const int oj = base + j;
        const real rx = s.px[oj] - mx, ry = s.py[oj] - my;
        const real q2 = rx * rx + ry * ry;
        const real sv = sx * rx + sy * ry;                                 // ww:67
        const bool ok = !((sv < (real)0) | (sv > p.range) | (q2 - sv * sv > p.r_p2));  // ww:68-69
        if (oj < eLo)        { if (ok && sv < bU) { bU = sv; iU = oj; } }
        else if (oj < eHi)   { if (ok && sv < bE) { bE = sv; iE = oj; } }
        else if (oj < p.Nall){ if (ok && sv < bP) { bP = sv; iP = oj; } }
        else                 { if (ok && sv < bO) { bO = sv; } }
      }
    }
    if (kval) {  // ww:312-353, 388-395: feature-major, sensor-minor
      const bool hO = bO < INF, hE = bE < INF, hP = bP < INF, hU = bU < INF;
      const real z = (real)0;
      if (p.speed_features) {
        const real sE = hE ? sx * (s.vx[iE] - mvx) + sy * (s.vy[iE] - mvy) : z;
        const real sP = hP ? sx * (s.vx[iP] - mvx) + sy * (s.vy[iP] - mvy) : z;
        const real sU = hU ? sx * (s.vx[iU] - mvx) + sy * (s.vy[iU] - mvy) : z;
        store_stream(obs_row + 0 * p.K + k, hO ? bO : z);
        store_stream(obs_row + 1 * p.K + k, hE ? bE : z);
        store_stream(obs_row + 2 * p.K + k, sE);
        store_stream(obs_row + 3 * p.K + k, hP ? bP : z);
        store_stream(obs_row + 4 * p.K + k, sP);
        store_stream(obs_row + 5 * p.K + k, hU ? bU : z);
        store_stream(obs_row + 6 * p.K + k, sU);
      } else {
        store_stream(obs_row + 0 * p.K + k, hO ? bO : z);
        store_stream(obs_row + 1 * p.K + k, hE ? bE : z);
        store_stream(obs_row + 2 * p.K + k, hP ? bP : z);
        store_stream(obs_row + 3 * p.K + k, hU ? bU : z);
      }
    }
  }
  // ww:411-428 tail: collided-with-evader, collided-with-poison, id
  const int tail = p.K * (p.speed_features ? 7 : 4);
  if (lane == 0) store_stream(obs_row + tail, anyE ? (real)1 : (real)0);
  if (lane == 1) store_stream(obs_row + tail + 1, anyP ? (real)1 : (real)0);
  if (lane == 2 && p.addid) store_stream(obs_row + tail + 2, (real)(pi + 1));
}

template <typename real>
__global__ void __launch_bounds__(512) ww_kernel(const __grid_constant__ WWParams<real> p) {
  extern __shared__ __align__(16) char smem[];
  real* sens = reinterpret_cast<real*>(smem);
  char* slots = smem + (((size_t)2 * p.K * sizeof(real) + 15) / 16) * 16;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
  const int e0 = (int)(((long long)blockIdx.x * p.E) / gridDim.x);
  const int e1 = (int)(((long long)(blockIdx.x + 1) * p.E) / gridDim.x);
  const int n_loc = e1 - e0;
  if (n_loc <= 0) return;

  // ---- stage state: HBM (SoA, env minor) -> shared memory --------------------------------
  for (int i = tid; i < 2 * p.K; i += blockDim.x) sens[i] = p.sensors[i];
  for (int i = tid; i < n_loc * p.Nall; i += blockDim.x) {
    const int el = i % n_loc, o = i / n_loc;
    const EnvSlot<real> s = env_slot(p, slots, el);
    const size_t g = (size_t)o * p.E + e0 + el;
    s.px[o] = p.pos_x[g]; s.py[o] = p.pos_y[g]; s.vx[o] = p.vel_x[g]; s.vy[o] = p.vel_y[g];
  }
  for (int el = tid; el < n_loc; el += blockDim.x) {
    const EnvSlot<real> s = env_slot(p, slots, el);
    s.px[p.Nall] = p.obst_px[e0 + el];
    s.py[p.Nall] = p.obst_py[e0 + el];
    s.meta[0] = p.timestep[e0 + el];
    s.meta[1] = p.path_len[e0 + el];
    s.meta[2] = (p.mode == 1) ? ((p.mask == nullptr || p.mask[e0 + el]) ? 1 : 0) : 0;
    *s.ctr = p.ctr[e0 + el];
  }
  __syncthreads();

  const int n_items = n_loc * p.Np;
  if (p.mode == 1) {
    // ---- reset(): draws + the reference's internal step(zeros), obs only (ww:144-172) ------
    for (int el = warp; el < n_loc; el += NW) {
      const EnvSlot<real> s = env_slot(p, slots, el);
      if (s.meta[2]) { env_reset_draws(p, s, e0 + el, lane); env_pre(p, s, e0 + el, 0, true, lane); }
    }
    __syncthreads();
    for (int it = warp; it < n_items; it += NW) {
      const int el = it / p.Np, pi = it - el * p.Np;
      const EnvSlot<real> s = env_slot(p, slots, el);
      if (s.meta[2]) sense_item(p, s, sens, pi, lane, p.obs + ((size_t)(e0 + el) * p.Np + pi) * p.D);
    }
    __syncthreads();
    for (int el = warp; el < n_loc; el += NW) {
      const EnvSlot<real> s = env_slot(p, slots, el);
      if (s.meta[2]) env_post(p, s, e0 + el, 0, true, lane);
    }
  } else {
    for (int t = 0; t < p.T; ++t) {
      for (int el = warp; el < n_loc; el += NW)
        env_pre(p, env_slot(p, slots, el), e0 + el, t, false, lane);
      __syncthreads();
      for (int it = warp; it < n_items; it += NW) {
        const int el = it / p.Np, pi = it - el * p.Np;
        sense_item(p, env_slot(p, slots, el), sens, pi, lane,
                   p.obs + (((size_t)t * p.E + e0 + el) * p.Np + pi) * p.D);
      }
      __syncthreads();
      int need = 0;
      for (int el = warp; el < n_loc; el += NW) {
        const EnvSlot<real> s = env_slot(p, slots, el);
        env_post(p, s, e0 + el, t, false, lane);
        need |= s.meta[2];
      }
      if (p.auto_reset) {
        // VecEnvExecutor.step: a done env is reset in place; its obs slot gets the reset obs
        // (rllab/sandbox/rocky/tf/envs/vec_env_executor.py:24-27).
        if (__syncthreads_or(need)) {
          for (int el = warp; el < n_loc; el += NW) {
            const EnvSlot<real> s = env_slot(p, slots, el);
            if (s.meta[2]) { env_reset_draws(p, s, e0 + el, lane); env_pre(p, s, e0 + el, t, true, lane); }
          }
          __syncthreads();
          for (int it = warp; it < n_items; it += NW) {
            const int el = it / p.Np, pi = it - el * p.Np;
            const EnvSlot<real> s = env_slot(p, slots, el);
            if (s.meta[2])
              sense_item(p, s, sens, pi, lane, p.obs + (((size_t)t * p.E + e0 + el) * p.Np + pi) * p.D);
          }
          __syncthreads();
          for (int el = warp; el < n_loc; el += NW) {
            const EnvSlot<real> s = env_slot(p, slots, el);
            if (s.meta[2]) env_post(p, s, e0 + el, t, true, lane);
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- write state back ---------------------------------------------------------------------
  for (int i = tid; i < n_loc * p.Nall; i += blockDim.x) {
    const int el = i % n_loc, o = i / n_loc;
    const EnvSlot<real> s = env_slot(p, slots, el);
    const size_t g = (size_t)o * p.E + e0 + el;
    p.pos_x[g] = s.px[o]; p.pos_y[g] = s.py[o]; p.vel_x[g] = s.vx[o]; p.vel_y[g] = s.vy[o];
  }
  for (int el = tid; el < n_loc; el += blockDim.x) {
    const EnvSlot<real> s = env_slot(p, slots, el);
    p.obst_px[e0 + el] = s.px[p.Nall];
    p.obst_py[e0 + el] = s.py[p.Nall];
    p.timestep[e0 + el] = s.meta[0];
    p.path_len[e0 + el] = s.meta[1];
    p.ctr[e0 + el] = *s.ctr;
  }
}

}  // namespace madrl

// =================================================================================================
// Host side: C ABI
// =================================================================================================
using namespace madrl;

struct madrl_ww {
  madrl_ww_config cfg;
  madrl_ww_layout lay;
  char* state;
  bool owns_state;
  int device, sms;
  int warps_per_block, blocks_per_sm;
  // staging for the host-buffer entry points (lazily sized)
  void* stage;
  size_t stage_bytes;
};

static int ww_validate(const madrl_ww_config* c) {
  MADRL_REQUIRE(c != nullptr, "config is NULL");
  MADRL_REQUIRE(c->n_envs > 0, "n_envs must be > 0");
  MADRL_REQUIRE(c->n_pursuers >= 1 && c->n_pursuers <= 32,
                "n_pursuers must be in [1,32] (collision rows are 32-bit masks), got %d", c->n_pursuers);
  MADRL_REQUIRE(c->n_evaders >= 1 && c->n_poison >= 1, "n_evaders and n_poison must be >= 1");
  MADRL_REQUIRE(c->n_sensors >= 1 && c->n_sensors <= 1024, "n_sensors out of range");
  MADRL_REQUIRE(c->n_coop >= 1, "n_coop must be >= 1");
  MADRL_REQUIRE(c->timestep_limit >= 1, "timestep_limit must be >= 1");
  return MADRL_OK;
}

extern "C" int madrl_ww_state_layout(const madrl_ww_config* c, madrl_ww_layout* out) {
  int rc = ww_validate(c);
  if (rc) return rc;
  MADRL_REQUIRE(out != nullptr, "layout out is NULL");
  const size_t rb = c->fp64 ? 8 : 4, E = (size_t)c->n_envs;
  const size_t nobj = (size_t)c->n_pursuers + c->n_evaders + c->n_poison;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  out->rng_counter = take(8 * E);
  out->pos_x = take(rb * nobj * E);
  out->pos_y = take(rb * nobj * E);
  out->vel_x = take(rb * nobj * E);
  out->vel_y = take(rb * nobj * E);
  out->obst_x = take(rb * E);
  out->obst_y = take(rb * E);
  out->timestep = take(4 * E);
  out->path_len = take(4 * E);
  out->sensors = take(rb * 2 * (size_t)c->n_sensors);
  out->total_bytes = off;
  out->n_obj = (int32_t)nobj;
  out->obs_dim = c->n_sensors * (c->speed_features ? 7 : 4) + 2 + (c->addid ? 1 : 0);  // ww:18-24
  out->real_bytes = (int32_t)rb;
  out->_pad = 0;
  return MADRL_OK;
}

template <typename real>
static int ww_upload_sensors(madrl_ww* h) {
  const int K = h->cfg.n_sensors;
  real* tab = new (std::nothrow) real[2 * (size_t)K];
  if (!tab) return MADRL_ENOMEM;
  // ww:29-31  angles = linspace(0, 2pi, K+1)[:-1]
  const double step = (2.0 * M_PI - 0.0) / (double)K;
  for (int k = 0; k < K; ++k) {
    const double a = (double)k * step + 0.0;
    tab[k] = (real)cos(a);
    tab[K + k] = (real)sin(a);
  }
  cudaError_t e = cudaMemcpy(h->state + h->lay.sensors, tab, sizeof(real) * 2 * K, cudaMemcpyHostToDevice);
  delete[] tab;
  MADRL_CUDA_CHECK(e);
  return MADRL_OK;
}

extern "C" int madrl_ww_create(const madrl_ww_config* c, void* state_dev, madrl_ww** out) {
  MADRL_REQUIRE(out != nullptr, "out is NULL");
  madrl_ww_layout lay;
  int rc = madrl_ww_state_layout(c, &lay);
  if (rc) return rc;
  madrl_ww* h = new (std::nothrow) madrl_ww();
  if (!h) return MADRL_ENOMEM;
  h->cfg = *c;
  h->lay = lay;
  h->stage = nullptr;
  h->stage_bytes = 0;
  h->warps_per_block = 0;
  h->blocks_per_sm = 0;
  cudaError_t e = cudaGetDevice(&h->device);
  if (e != cudaSuccess) { set_error("cudaGetDevice: %s", cudaGetErrorString(e)); delete h; return MADRL_ECUDA; }
  h->sms = sm_count(h->device);
  if (h->sms <= 0) { delete h; return MADRL_ECUDA; }
  if (state_dev) {
    h->state = (char*)state_dev;
    h->owns_state = false;
  } else {
    e = cudaMalloc((void**)&h->state, lay.total_bytes);
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu): %s", lay.total_bytes, cudaGetErrorString(e)); delete h; return MADRL_ENOMEM; }
    h->owns_state = true;
  }
  e = cudaMemset(h->state, 0, lay.total_bytes);
  if (e != cudaSuccess) { set_error("cudaMemset: %s", cudaGetErrorString(e)); madrl_ww_destroy(h); return MADRL_ECUDA; }
  rc = c->fp64 ? ww_upload_sensors<double>(h) : ww_upload_sensors<float>(h);
  if (rc) { madrl_ww_destroy(h); return rc; }
  *out = h;
  return MADRL_OK;
}

extern "C" int madrl_ww_destroy(madrl_ww* h) {
  if (!h) return MADRL_OK;
  if (h->owns_state && h->state) cudaFree(h->state);
  if (h->stage) cudaFree(h->stage);
  delete h;
  return MADRL_OK;
}

extern "C" void* madrl_ww_state_ptr(madrl_ww* h) { return h ? h->state : nullptr; }

extern "C" int madrl_ww_seed(madrl_ww* h, uint64_t seed, void* stream) {
  MADRL_REQUIRE(h != nullptr, "handle is NULL");
  h->cfg.seed = seed;
  MADRL_CUDA_CHECK(cudaMemsetAsync(h->state + h->lay.rng_counter, 0, 8 * (size_t)h->cfg.n_envs,
                                   (cudaStream_t)stream));
  return MADRL_OK;
}

extern "C" int madrl_ww_set_launch(madrl_ww* h, int warps_per_block, int blocks_per_sm) {
  MADRL_REQUIRE(h != nullptr, "handle is NULL");
  MADRL_REQUIRE(warps_per_block >= 0 && warps_per_block <= 16, "warps_per_block must be in [0,16]");
  MADRL_REQUIRE(blocks_per_sm >= 0 && blocks_per_sm <= 32, "blocks_per_sm must be in [0,32]");
  h->warps_per_block = warps_per_block;
  h->blocks_per_sm = blocks_per_sm;
  return MADRL_OK;
}

template <typename real>
static int ww_launch(madrl_ww* h, int mode, int T, const void* actions, void* obs, void* rew,
                     uint8_t* done, int32_t* info, const uint8_t* mask, int auto_reset,
                     cudaStream_t stream) {
  const madrl_ww_config& c = h->cfg;
  WWParams<real> p;
  p.E = c.n_envs; p.env_id_base = c.env_id_base;
  p.Np = c.n_pursuers; p.Ne = c.n_evaders; p.Npo = c.n_poison; p.K = c.n_sensors;
  p.n_coop = c.n_coop; p.D = h->lay.obs_dim; p.Nall = h->lay.n_obj;
  p.CW = (p.Nall + 1 + 31) / 32;
  p.reward_global = c.reward_global; p.addid = c.addid; p.speed_features = c.speed_features;
  p.random_obstacle = c.random_obstacle; p.timestep_limit = c.timestep_limit;
  p.max_path_length = c.max_path_length;
  p.T = T; p.mode = mode; p.auto_reset = auto_reset;
  // radii ww:108-118 (double arithmetic as in the reference, then narrowed once)
  const double r_p = c.radius, r_e = c.radius * 2, r_po = c.radius * 3 / 4;
  p.r_p2 = (real)(r_p * r_p);
  p.range = (real)c.sensor_range;
  // Exact conservative cull: sv <= range and d2 - sv^2 <= r^2 imply d2 <= range^2 + r^2.
  p.cull2 = (real)((c.sensor_range * c.sensor_range + r_p * r_p) * (1.0 + 1e-4) + 1e-12);
  p.rsum_e = (real)(r_p + r_e);
  p.rsum_po = (real)(r_p + r_po);
  p.thr_obst_p = (real)(r_p + c.obstacle_radius);
  p.thr_obst_e = (real)(r_e + c.obstacle_radius);
  p.thr_obst_po = (real)(r_po + c.obstacle_radius);
  p.thr_resp_p = (real)(r_p * 2 + c.obstacle_radius);
  p.thr_resp_e = (real)(r_e * 2 + c.obstacle_radius);
  p.thr_resp_po = (real)(r_po * 2 + c.obstacle_radius);
  p.obst_x = (real)c.obstacle_x; p.obst_y = (real)c.obstacle_y;
  p.ev_speed = (real)c.ev_speed; p.poison_speed = (real)c.poison_speed;
  p.action_scale = (real)c.action_scale;
  p.poison_reward = (real)c.poison_reward; p.food_reward = (real)c.food_reward;
  p.encounter_reward = (real)c.encounter_reward; p.control_penalty = (real)c.control_penalty;
  p.seed = c.seed;
  char* st = h->state;
  p.pos_x = (real*)(st + h->lay.pos_x); p.pos_y = (real*)(st + h->lay.pos_y);
  p.vel_x = (real*)(st + h->lay.vel_x); p.vel_y = (real*)(st + h->lay.vel_y);
  p.obst_px = (real*)(st + h->lay.obst_x); p.obst_py = (real*)(st + h->lay.obst_y);
  p.timestep = (int32_t*)(st + h->lay.timestep); p.path_len = (int32_t*)(st + h->lay.path_len);
  p.ctr = (uint64_t*)(st + h->lay.rng_counter);
  p.sensors = (const real*)(st + h->lay.sensors);
  p.actions = (const real*)actions; p.obs = (real*)obs; p.rew = (real*)rew;
  p.done = done; p.info = info; p.mask = mask;

  // launch geometry: persistent blocks, contiguous env ranges, SM-balanced
  const int wpb = h->warps_per_block > 0 ? h->warps_per_block : 8;
  const int bps = h->blocks_per_sm > 0 ? h->blocks_per_sm : 8;
  int grid = h->sms * bps;
  if (grid > p.E) grid = p.E;
  p.max_loc = (p.E + grid - 1) / grid;
  // smem slot layout
  size_t off = 0;
  const size_t rb = sizeof(real);
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 8); return o; };
  take(rb * (p.Nall + 1));                        // px (+ obstacle)
  p.off_py = (int)take(rb * (p.Nall + 1));
  p.off_vx = (int)take(rb * p.Nall);
  p.off_vy = (int)take(rb * p.Nall);
  p.off_rew = (int)take(rb * p.Np);
  p.off_coll = (int)take(4 * (size_t)p.Np * p.CW);
  p.off_meta = (int)take(4 * 4);
  p.off_ctr = (int)take(8);
  p.env_stride = (int)align_up(off, 16);
  const size_t smem = align_up(2 * (size_t)p.K * rb, 16) + (size_t)p.max_loc * p.env_stride;
  MADRL_REQUIRE(smem <= 200 * 1024, "env batch per block needs %zu B of shared memory", smem);
  if (smem > 48 * 1024) {
    MADRL_CUDA_CHECK(cudaFuncSetAttribute(ww_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  ww_kernel<real><<<grid, wpb * 32, smem, stream>>>(p);
  g_launches.fetch_add(1);
  MADRL_CUDA_CHECK(cudaGetLastError());
  return MADRL_OK;
}

extern "C" int madrl_ww_reset(madrl_ww* h, const uint8_t* mask_dev, void* obs_dev, void* stream) {
  MADRL_REQUIRE(h != nullptr && obs_dev != nullptr, "handle/obs is NULL");
  return h->cfg.fp64 ? ww_launch<double>(h, 1, 1, nullptr, obs_dev, nullptr, nullptr, nullptr, mask_dev, 0, (cudaStream_t)stream)
                     : ww_launch<float>(h, 1, 1, nullptr, obs_dev, nullptr, nullptr, nullptr, mask_dev, 0, (cudaStream_t)stream);
}

extern "C" int madrl_ww_rollout(madrl_ww* h, int T, const void* actions_dev, void* obs_dev,
                                void* rew_dev, uint8_t* done_dev, int32_t* info_dev,
                                int auto_reset, void* stream) {
  MADRL_REQUIRE(h != nullptr, "handle is NULL");
  MADRL_REQUIRE(T >= 1, "T must be >= 1");
  MADRL_REQUIRE(actions_dev && obs_dev && rew_dev && done_dev && info_dev, "NULL trajectory buffer");
  return h->cfg.fp64 ? ww_launch<double>(h, 0, T, actions_dev, obs_dev, rew_dev, done_dev, info_dev, nullptr, auto_reset, (cudaStream_t)stream)
                     : ww_launch<float>(h, 0, T, actions_dev, obs_dev, rew_dev, done_dev, info_dev, nullptr, auto_reset, (cudaStream_t)stream);
}

extern "C" int madrl_ww_step(madrl_ww* h, const void* actions_dev, void* obs_dev, void* rew_dev,
                             uint8_t* done_dev, int32_t* info_dev, int auto_reset, void* stream) {
  return madrl_ww_rollout(h, 1, actions_dev, obs_dev, rew_dev, done_dev, info_dev, auto_reset, stream);
}

// ---- host-buffer entry points -------------------------------------------------------------------
static int ww_stage(madrl_ww* h, size_t bytes) {
  if (h->stage_bytes >= bytes) return MADRL_OK;
  if (h->stage) cudaFree(h->stage);
  h->stage = nullptr;
  h->stage_bytes = 0;
  cudaError_t e = cudaMalloc(&h->stage, bytes);
  if (e != cudaSuccess) { set_error("cudaMalloc(stage %zu): %s", bytes, cudaGetErrorString(e)); return MADRL_ENOMEM; }
  h->stage_bytes = bytes;
  return MADRL_OK;
}

extern "C" int madrl_ww_reset_host(madrl_ww* h, const uint8_t* mask_host, void* obs_host) {
  MADRL_REQUIRE(h != nullptr && obs_host != nullptr, "handle/obs is NULL");
  const size_t E = h->cfg.n_envs, rb = h->lay.real_bytes;
  const size_t obs_b = E * h->cfg.n_pursuers * h->lay.obs_dim * rb;
  const size_t mask_off = align_up(obs_b, 256);
  int rc = ww_stage(h, mask_off + E);
  if (rc) return rc;
  char* st = (char*)h->stage;
  uint8_t* mask_dev = nullptr;
  if (mask_host) {
    mask_dev = (uint8_t*)(st + mask_off);
    MADRL_CUDA_CHECK(cudaMemcpyAsync(mask_dev, mask_host, E, cudaMemcpyHostToDevice, 0));
    // rows of unmasked envs must survive: seed the staging buffer with the caller's obs
    MADRL_CUDA_CHECK(cudaMemcpyAsync(st, obs_host, obs_b, cudaMemcpyHostToDevice, 0));
  }
  rc = madrl_ww_reset(h, mask_dev, st, nullptr);
  if (rc) return rc;
  MADRL_CUDA_CHECK(cudaMemcpyAsync(obs_host, st, obs_b, cudaMemcpyDeviceToHost, 0));
  MADRL_CUDA_CHECK(cudaStreamSynchronize(0));
  return MADRL_OK;
}

extern "C" int madrl_ww_rollout_host(madrl_ww* h, int T, const void* actions_host, void* obs_host,
                                     void* rew_host, uint8_t* done_host, int32_t* info_host,
                                     int auto_reset) {
  MADRL_REQUIRE(h != nullptr, "handle is NULL");
  MADRL_REQUIRE(T >= 1, "T must be >= 1");
  MADRL_REQUIRE(actions_host && obs_host && rew_host && done_host && info_host, "NULL trajectory buffer");
  const size_t E = h->cfg.n_envs, Np = h->cfg.n_pursuers, rb = h->lay.real_bytes, TT = (size_t)T;
  const size_t act_b = TT * E * Np * 2 * rb, obs_b = TT * E * Np * h->lay.obs_dim * rb;
  const size_t rew_b = TT * E * Np * rb, done_b = TT * E, info_b = TT * E * 2 * 4;
  const size_t o_act = 0, o_obs = align_up(o_act + act_b, 256), o_rew = align_up(o_obs + obs_b, 256);
  const size_t o_done = align_up(o_rew + rew_b, 256), o_info = align_up(o_done + done_b, 256);
  int rc = ww_stage(h, o_info + info_b);
  if (rc) return rc;
  char* st = (char*)h->stage;
  MADRL_CUDA_CHECK(cudaMemcpyAsync(st + o_act, actions_host, act_b, cudaMemcpyHostToDevice, 0));
  rc = madrl_ww_rollout(h, T, st + o_act, st + o_obs, st + o_rew, (uint8_t*)(st + o_done),
                        (int32_t*)(st + o_info), auto_reset, nullptr);
  if (rc) return rc;
  MADRL_CUDA_CHECK(cudaMemcpyAsync(obs_host, st + o_obs, obs_b, cudaMemcpyDeviceToHost, 0));
  MADRL_CUDA_CHECK(cudaMemcpyAsync(rew_host, st + o_rew, rew_b, cudaMemcpyDeviceToHost, 0));
  MADRL_CUDA_CHECK(cudaMemcpyAsync(done_host, st + o_done, done_b, cudaMemcpyDeviceToHost, 0));
  MADRL_CUDA_CHECK(cudaMemcpyAsync(info_host, st + o_info, info_b, cudaMemcpyDeviceToHost, 0));
  MADRL_CUDA_CHECK(cudaStreamSynchronize(0));
  return MADRL_OK;
}
