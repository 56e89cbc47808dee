/* madrl_b200 -- C ABI of the B200 batched multi-agent environment engine.
 *
 * One handle = E independent environment instances of one family, resident in HBM (one
 * struct-of-arrays record per env) and stepped in lockstep by sm_100a CUDA kernels.  Every entry
 * point replaces, for a whole batch, one method of the reference's Python environment interface
 * (`madrl_environments/__init__.py:27-119`, AbstractMAEnv):
 *
 *   madrl_ww_*      <-> MAWaterWorld            madrl_environments/pursuit/waterworld.py
 *        _reset     <-> reset()                 :144-172
 *        _step      <-> step(action)            :220-436   (T = 1)
 *        _rollout   <-> T x step() under rllab's VecEnvExecutor.step auto-reset contract
 *                       rllab/sandbox/rocky/tf/envs/vec_env_executor.py:16-28
 *        _seed      <-> seed(seed)              :135-137
 *   madrl_pursuit_* <-> PursuitEvade            madrl_environments/pursuit/pursuit_evade.py
 *        _reset :173-207, _step :209-262, _seed :166-168
 *   madrl_hostage_* <-> ContinuousHostageWorld  madrl_environments/hostage.py
 *        _reset :142-179, _step :228-429, _seed :134-136
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success or a negative MADRL_E* code and
 *     never throws; `madrl_last_error()` returns a thread-local message for the last failure.
 *   - `*_dev` pointers are device pointers owned by the caller (e.g. torch tensors); the
 *     `_host` variants take host pointers and perform the host<->device copies themselves.
 *   - the per-env state lives in one device blob.  The caller may provide it
 *     (`state_dev`, size from `*_state_layout`) or pass NULL to let the library cudaMalloc it.
 *     The layout (byte offsets of the per-env records and constant tables) is reported by
 *     `*_state_layout` so that parity harnesses and `set_param_values`-style callers can read
 *     and write state directly.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  One handle is used
 *     from one host thread / stream at a time.
 *   - randomness: counter-based Philox4x32-10 streams keyed by (seed, global env id); the
 *     per-env draw counter is part of the state (csrc/philox.cuh).
 *   - `real` below means float when cfg.fp64 == 0 (production) and double when cfg.fp64 == 1
 *     (verification build used by the parity tests).
 */
#ifndef MADRL_B200_H
#define MADRL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MADRL_OK 0
#define MADRL_EINVAL (-1)   /* bad argument / unsupported configuration */
#define MADRL_ECUDA (-2)    /* CUDA runtime error (see madrl_last_error) */
#define MADRL_ENOMEM (-3)

const char* madrl_last_error(void);
int madrl_version(void);
/* sizeof(madrl_{ww,pursuit,hostage}_{config,layout}) as this library was compiled: a foreign-language
 * binding compares them with its own struct definitions before the first call. */
void madrl_abi_sizes(int32_t* out6);
/* Number of kernels launched by this library since load (bench.py's gpu_launches). */
uint64_t madrl_launch_count(void);
/* Device->host bytes per pipeline chunk of the `*_rollout_host` entry points (0 = default 32 MB). */
void madrl_set_host_chunk_bytes(size_t bytes);

/* CUDA-IPC buffers for the fused multi-GPU exchange (see madrl_ww_set_peers): allocate + export on
 * the owning rank, open on every other rank with ITS device current (peer access over NVLink is
 * enabled lazily), close / free at teardown.  handle64 is the 64-byte cudaIpcMemHandle_t. */
int madrl_ipc_alloc(size_t bytes, void** ptr, unsigned char* handle64);
int madrl_ipc_open(const unsigned char* handle64, void** ptr);
int madrl_ipc_close(void* ptr);
int madrl_ipc_free(void* ptr);

/* Stream-ordered primitives of the per-rollout multi-GPU exchange (madrl_b200/dist.py
 * AsyncRootGather): executed by the copy engines / the stream front end, no SM involved, so the
 * persistent rollout kernel of the next rollout is not disturbed.  The reference's counterpart is
 * the workers' pickled path return to the master (rllab/rllab/sampler/stateful_pool.py:102-157).
 *   madrl_copy_async          device->device copy; either side may be a CUDA-IPC peer mapping (NVLink)
 *   madrl_stream_write32      *addr = value when the stream reaches this point (LOCAL device memory)
 *   madrl_stream_wait_geq32   the stream waits until (int32)(*addr - value) >= 0 (LOCAL device memory)
 *   madrl_stream_memops_available   1 if the driver exports the two stream memory operations */
int madrl_stream_memops_available(void);
int madrl_copy_async(void* dst_dev, const void* src_dev, size_t bytes, void* stream);
int madrl_stream_write32(void* stream, void* addr_dev, uint32_t value);
int madrl_stream_wait_geq32(void* stream, void* addr_dev, uint32_t value);

/* ------------------------------------------------------------------ MAWaterWorld ------------ */
typedef struct madrl_ww_config {
  int32_t n_envs;          /* envs in THIS handle (the local shard)                              */
  int32_t env_id_base;     /* global id of local env 0: RNG key is (seed, env_id_base + e)       */
  int32_t n_pursuers, n_evaders, n_poison, n_sensors, n_coop;   /* waterworld.py:77-79          */
  int32_t reward_global;   /* reward_mech == 'global'                                            */
  int32_t addid, speed_features;
  int32_t random_obstacle; /* obstacle_loc is None -> drawn at reset (waterworld.py:147-148)     */
  int32_t timestep_limit;  /* 1000 (waterworld.py:124-126)                                       */
  int32_t max_path_length; /* VecEnvExecutor horizon, 0 = none                                   */
  int32_t fp64;            /* 0: fp32 arithmetic and buffers; 1: fp64 verification build         */
  double radius, obstacle_radius, obstacle_x, obstacle_y, ev_speed, poison_speed, sensor_range,
      action_scale, poison_reward, food_reward, encounter_reward, control_penalty;
  uint64_t seed;
} madrl_ww_config;

/* Byte offsets into the state blob.  The dynamic state is one record per env:
 * objs real [E][4][n_obj] = rows x, y, vx, vy; objects ordered pursuers, evaders, poisons
 * (one warp owns one env, so a record is one coalesced read per launch). */
typedef struct madrl_ww_layout {
  size_t total_bytes;
  size_t objs;                         /* real [E][4][n_obj]                                   */
  size_t obst;                         /* real [E][2] obstacle centre                          */
  size_t timestep;                     /* int32 [E]  env._timesteps                            */
  size_t path_len;                     /* int32 [E]  VecEnvExecutor.ts                         */
  size_t rng_counter;                  /* uint64 [E] draws consumed                            */
  size_t sensors;                      /* real [2][n_sensors] unit vectors (cos row, sin row)  */
  int32_t n_obj, obs_dim, real_bytes, _pad;
} madrl_ww_layout;

typedef struct madrl_ww madrl_ww;

int madrl_ww_state_layout(const madrl_ww_config* cfg, madrl_ww_layout* out);
int madrl_ww_create(const madrl_ww_config* cfg, void* state_dev, madrl_ww** out);
int madrl_ww_destroy(madrl_ww* h);
void* madrl_ww_state_ptr(madrl_ww* h);
/* seed(): new key, draw counters reset to 0 (waterworld.py:135-137 creates a fresh generator). */
int madrl_ww_seed(madrl_ww* h, uint64_t seed, void* stream);
/* Launch geometry override (0 = library default): resident blocks per SM.  Blocks are one warp (= one
 * env) since round 2; `warps_per_block` is accepted for ABI stability and ignored. */
int madrl_ww_set_launch(madrl_ww* h, int warps_per_block, int blocks_per_sm);
/* Terminal observations: with auto_reset the obs slot of a done step holds the RESET observation
 * (VecEnvExecutor semantics); StandardizedEnv (madrl_environments/__init__.py:283-291) also sees the
 * terminal one.  After this call every `*_rollout` (device pointers) with auto_reset first copies the rows
 * of a finished env to the same [t][e] slot of term_obs_dev (same shape as that rollout's obs tensor;
 * other slots are left untouched).  NULL switches it off. */
int madrl_ww_set_terminal_obs(madrl_ww* h, void* term_obs_dev);
/* Fused per-rollout exchange for env-sharded multi-GPU runs: after this call every rollout also
 * stores its reward / done / info rows into slot `slot` of each of the `n_dest` listed gather
 * buffers (peer-mapped device pointers: CUDA-IPC mappings of other ranks' buffers over NVLink; one
 * destination = gather-to-root, world destinations = all-gather).  Env-major layout per buffer:
 *   rew  real  [n_slots][E][t_max][Np]    done uint8 [n_slots][E][t_max]
 *   info int32 [n_slots][E][t_max][2]
 * Rows are staged in registers and written as coalesced runs.  A stream-ordered barrier between the
 * ranks after the rollout completes the gather.  n_dest = 0 disables. */
int madrl_ww_set_peers(madrl_ww* h, int n_dest, int slot, int t_max, void* const* rew_dest,
                       void* const* done_dest, void* const* info_dest);

/* reset(): envs with mask_dev[e] != 0 (all if NULL) are re-initialised and advanced by the
 * reference's internal step(zeros); obs_dev real [E][Np][obs_dim], rows of unmasked envs untouched. */
int madrl_ww_reset(madrl_ww* h, const uint8_t* mask_dev, void* obs_dev, void* stream);
/* T lockstep steps in ONE launch.  actions_dev real [T][E][Np][2]; obs_dev real [T][E][Np][obs_dim];
 * rew_dev real [T][E][Np]; done_dev uint8 [T][E]; info_dev int32 [T][E][2] = (evcatches, pocatches),
 * 8-byte aligned (a row is one 8-byte store).
 * auto_reset != 0: VecEnvExecutor.step semantics -- a done env is reset in place and the obs
 * slot of that step holds the reset observation. */
int madrl_ww_rollout(madrl_ww* h, int T, const void* actions_dev, void* obs_dev, void* rew_dev,
                     uint8_t* done_dev, int32_t* info_dev, int auto_reset, void* stream);
/* Closed-loop rollout: the actions come from the reference's hand-written policy
 * (heuristics/waterworld.py:11-53 WaterworldHeuristicPolicy.sample_actions: flee the obstacle and poison,
 * chase evaders, close in on allies; unit vector or zero), evaluated INSIDE the rollout kernel on the
 * features each warp has just computed -- no action tensor is read, no per-step launch.
 * obs0_dev real [E][Np][obs_dim]: the observation the first action is computed from (the reset
 * observation, or the last observation of the previous rollout); actions_out_dev real [T][E][Np][2]
 * receives the actions taken (NULL = not recorded).  Needs speed_features (the 7K layout the policy
 * indexes).  Each agent row is normalised by its own norm (the reference is called per agent, B = 1).
 * Other arguments as madrl_ww_rollout. */
int madrl_ww_rollout_heuristic(madrl_ww* h, int T, const void* obs0_dev, void* actions_out_dev,
                               void* obs_dev, void* rew_dev, uint8_t* done_dev, int32_t* info_dev,
                               int auto_reset, void* stream);
/* step() == rollout with T = 1. */
int madrl_ww_step(madrl_ww* h, const void* actions_dev, void* obs_dev, void* rew_dev,
                  uint8_t* done_dev, int32_t* info_dev, int auto_reset, void* stream);
/* Host-buffer variants (what a non-torch FFI caller binds): copies are part of the call and the
 * call returns after the results are in the host buffers (pinned memory for full PCIe speed).  The
 * rollout is chunked: the copy engines drain chunk c while chunk c+1 is computed.  They run on two
 * internal streams that are ordered after earlier work on the legacy default stream; work the caller
 * queued on other streams must be synchronised by the caller.
 * `_host2` takes flags: MADRL_HOST_OBS_LAST = only the LAST step's observations come back (obs_host is
 * then [E][Np][D]) -- the policy-on-device mode: rewards / dones / infos of every step + the
 * observation needed to continue; the call then runs at the kernel's rate instead of PCIe's. */
#define MADRL_HOST_OBS_LAST 1
int madrl_ww_reset_host(madrl_ww* h, const uint8_t* mask_host, void* obs_host);
int madrl_ww_rollout_host(madrl_ww* h, int T, const void* actions_host, void* obs_host,
                          void* rew_host, uint8_t* done_host, int32_t* info_host, int auto_reset);
int madrl_ww_rollout_host2(madrl_ww* h, int T, const void* actions_host, void* obs_host,
                           void* rew_host, uint8_t* done_host, int32_t* info_host, int auto_reset,
                           int flags);

/* ------------------------------------------------------------------ PursuitEvade ------------ */
typedef struct madrl_pursuit_config {
  int32_t n_envs, env_id_base;
  int32_t n_pursuers, n_evaders;      /* pursuit_evade.py:60-61 (<= 32 / <= 64)                 */
  int32_t xs, ys, n_maps;             /* map_pool shape (n_maps, xs, ys); -1 = building         */
  int32_t obs_range;                  /* pursuit_evade.py:63                                     */
  int32_t flatten;                    /* pursuit_evade.py:67; 0 = conv layout (R, R, 4)          */
  int32_t n_catch, surround;          /* pursuit_evade.py:79,142                                 */
  int32_t reward_global, include_id, sample_maps;
  int32_t max_path_length;            /* VecEnvExecutor horizon, 0 = none                        */
  int32_t max_opponents;              /* random_opponents (pursuit_evade.py:81-82,177-181): > 0 = every reset
                                         draws randint(1, max_opponents) live evaders (<= n_evaders); 0 = off */
  double layer_norm, catchr, term_pursuit, urgency_reward, constraint_window;
  uint64_t seed;
} madrl_pursuit_config;

/* Byte offsets into the state blob (one record per env). */
typedef struct madrl_pursuit_layout {
  size_t total_bytes;
  size_t pos;          /* uint8  [E][2][n_agents]  x row, y row; pursuers first, then evaders     */
  size_t gone;         /* uint64 [E]  bit j: evader j has been removed (evaders_gone)             */
  size_t map_id;       /* int32  [E]  index into the map pool                                     */
  size_t path_len;     /* int32  [E]  VecEnvExecutor.ts                                           */
  size_t rng_counter;  /* uint64 [E]                                                              */
  size_t stale;        /* uint16 [E][Np][R*R]  the never-cleared channels 1-2 of local_obs
                          (pursuit_evade.py:119,438): pursuer count | evader count << 8           */
  size_t maps;         /* uint32 [n_maps][(xs+2p)(ys+2p)], p = (obs_range-1)/2+1: per map the empty cell grid
                          with its border of p marker cells, as the kernel keeps it in shared memory:
                          bit 0 building, bits 24-26 need_to_surround (pursuit_evade.py:523-540), border
                          cells 0x80000001 (constant)                                             */
  size_t lut, idv;     /* float tables (constant)                                                 */
  int32_t n_agents, obs_dim;
} madrl_pursuit_layout;

typedef struct madrl_pursuit madrl_pursuit;

int madrl_pursuit_state_layout(const madrl_pursuit_config* cfg, madrl_pursuit_layout* out);
/* map_pool_host: int32 [n_maps][xs][ys] (the .npy the reference loads, maps/map_pool16.npy). */
int madrl_pursuit_create(const madrl_pursuit_config* cfg, const int32_t* map_pool_host,
                         void* state_dev, madrl_pursuit** out);
int madrl_pursuit_destroy(madrl_pursuit* h);
void* madrl_pursuit_state_ptr(madrl_pursuit* h);
int madrl_pursuit_seed(madrl_pursuit* h, uint64_t seed, void* stream);
int madrl_pursuit_set_launch(madrl_pursuit* h, int warps_per_block, int blocks_per_sm);
/* Terminal observations: with auto_reset the obs slot of a done step holds the RESET observation
 * (VecEnvExecutor semantics); StandardizedEnv (madrl_environments/__init__.py:283-291) also sees the
 * terminal one.  After this call every `*_rollout` (device pointers) with auto_reset first copies the rows
 * of a finished env to the same [t][e] slot of term_obs_dev (same shape as that rollout's obs tensor;
 * other slots are left untouched).  NULL switches it off. */
int madrl_pursuit_set_terminal_obs(madrl_pursuit* h, void* term_obs_dev);
/* Curriculum knobs that survive pickling in the reference (pursuit_evade.py:264-272,397-411). */
int madrl_pursuit_set_params(madrl_pursuit* h, double catchr, double constraint_window);
/* reset(): obs_dev float [E][Np][obs_dim]. */
int madrl_pursuit_reset(madrl_pursuit* h, const uint8_t* mask_dev, float* obs_dev, void* stream);
/* actions_dev int32 [T][E][Np] in {0 left,1 right,2 up,3 down,4 stay} (DiscreteAgent.py:28-38);
 * obs_dev float [T][E][Np][obs_dim] (obs_dim = 3R^2+id if flatten else 4R^2 laid out [x][y][ch]);
 * rew_dev float [T][E][Np] (computed in float64, narrowed once);
 * done_dev uint8 [T][E]; info_dev int32 [T][E] = removed. */
int madrl_pursuit_rollout(madrl_pursuit* h, int T, const int32_t* actions_dev, float* obs_dev,
                          float* rew_dev, uint8_t* done_dev, int32_t* info_dev, int auto_reset,
                          void* stream);
/* Closed-loop rollout: the pursuers' actions come from the reference's hand-written policy
 * (heuristics/pursuit.py:18-50 PursuitHeuristicPolicy.sample_actions: walk towards the nearest evader of the
 * observation window, a random move when none is visible), evaluated INSIDE the rollout kernel on the window
 * each warp has just assembled.  obs0_dev float [E][Np][obs_dim]: the observation the first action is
 * computed from; actions_out_dev int32 [T][E][Np] receives the actions taken (NULL = not recorded).
 * floor_centre != 0: the window centre is (R//2, R//2) -- what `xs / 2` (heuristics/pursuit.py:23) gives under
 * Python 2, the reference's language; 0: (R/2, R/2) true division, what the same line gives under Python 3.
 * The policy's own draws (`action_space.sample()`, :48,50) are an injected counter-based stream: pursuer q
 * deciding on an observation produced when the env's draw counter stood at c takes word 32 c + q of the
 * (seed, env id, tag 2) Philox stream, mapped to {0..4} like every other range draw.  Other arguments as
 * madrl_pursuit_rollout. */
int madrl_pursuit_rollout_heuristic(madrl_pursuit* h, int T, const float* obs0_dev, int32_t* actions_out_dev,
                                    float* obs_dev, float* rew_dev, uint8_t* done_dev, int32_t* info_dev,
                                    int auto_reset, int floor_centre, void* stream);
int madrl_pursuit_step(madrl_pursuit* h, const int32_t* actions_dev, float* obs_dev, float* rew_dev,
                       uint8_t* done_dev, int32_t* info_dev, int auto_reset, void* stream);
int madrl_pursuit_reset_host(madrl_pursuit* h, const uint8_t* mask_host, float* obs_host);
int madrl_pursuit_rollout_host(madrl_pursuit* h, int T, const int32_t* actions_host, float* obs_host,
                               float* rew_host, uint8_t* done_host, int32_t* info_host,
                               int auto_reset);
int madrl_pursuit_rollout_host2(madrl_pursuit* h, int T, const int32_t* actions_host, float* obs_host,
                                float* rew_host, uint8_t* done_host, int32_t* info_host,
                                int auto_reset, int flags);

/* ------------------------------------------------------------------ ContinuousHostageWorld -- */
typedef struct madrl_hostage_config {
  int32_t n_envs, env_id_base;
  int32_t n_good, n_hostages, n_bad, n_coop_save, n_coop_avoid, n_sensors;   /* hostage.py:75-76 */
  int32_t reward_global, addid;
  int32_t random_key;       /* key_loc is None: drawn at the first reset, then kept (hostage.py:148) */
  int32_t timestep_limit, max_path_length, fp64;
  double radius, key_x, key_y, bad_speed, sensor_range, action_scale, save_reward, hit_reward,
      encounter_reward, not_saved_reward, bomb_reward, bomb_radius, key_radius, control_penalty;
  uint64_t seed;
} madrl_hostage_config;

typedef struct madrl_hostage_layout {
  size_t total_bytes;
  size_t objs;         /* real  [E][4][n_obj] rows x, y, vx, vy; rescuers, criminals, hostages   */
  size_t fixed;        /* real  [E][4] key x, key y, bomb x, bomb y                              */
  size_t saved;        /* uint8 [E][n_hostages] curr_host_saved_mask                             */
  size_t flags;        /* int32 [E] bit0 gate open, bit1 bombed, bit2 key location drawn         */
  size_t timestep, path_len;   /* int32 [E]                                                      */
  size_t rng_counter;  /* uint64 [E]                                                             */
  size_t sensors;      /* real [2][n_sensors]                                                    */
  int32_t n_obj, obs_dim, real_bytes, _pad;
} madrl_hostage_layout;

typedef struct madrl_hostage madrl_hostage;

int madrl_hostage_state_layout(const madrl_hostage_config* cfg, madrl_hostage_layout* out);
int madrl_hostage_create(const madrl_hostage_config* cfg, void* state_dev, madrl_hostage** out);
int madrl_hostage_destroy(madrl_hostage* h);
void* madrl_hostage_state_ptr(madrl_hostage* h);
int madrl_hostage_seed(madrl_hostage* h, uint64_t seed, void* stream);
int madrl_hostage_set_launch(madrl_hostage* h, int warps_per_block, int blocks_per_sm);
/* Terminal observations: with auto_reset the obs slot of a done step holds the RESET observation
 * (VecEnvExecutor semantics); StandardizedEnv (madrl_environments/__init__.py:283-291) also sees the
 * terminal one.  After this call every `*_rollout` (device pointers) with auto_reset first copies the rows
 * of a finished env to the same [t][e] slot of term_obs_dev (same shape as that rollout's obs tensor;
 * other slots are left untouched).  NULL switches it off. */
int madrl_hostage_set_terminal_obs(madrl_hostage* h, void* term_obs_dev);
/* obs_dev real [E][n_good][obs_dim] */
int madrl_hostage_reset(madrl_hostage* h, const uint8_t* mask_dev, void* obs_dev, void* stream);
/* actions_dev real [T][E][n_good][2]; obs_dev real [T][E][n_good][obs_dim]; rew_dev real
 * [T][E][n_good]; done_dev uint8 [T][E]; info_dev int32 [T][E][2] = (ho_saved, cr_encs), 8-byte
 * aligned (a row is one 8-byte store). */
int madrl_hostage_rollout(madrl_hostage* h, int T, const void* actions_dev, void* obs_dev,
                          void* rew_dev, uint8_t* done_dev, int32_t* info_dev, int auto_reset,
                          void* stream);
int madrl_hostage_step(madrl_hostage* h, const void* actions_dev, void* obs_dev, void* rew_dev,
                       uint8_t* done_dev, int32_t* info_dev, int auto_reset, void* stream);
int madrl_hostage_reset_host(madrl_hostage* h, const uint8_t* mask_host, void* obs_host);
int madrl_hostage_rollout_host(madrl_hostage* h, int T, const void* actions_host, void* obs_host,
                               void* rew_host, uint8_t* done_host, int32_t* info_host,
                               int auto_reset);
int madrl_hostage_rollout_host2(madrl_hostage* h, int T, const void* actions_host, void* obs_host,
                                void* rew_host, uint8_t* done_host, int32_t* info_host,
                                int auto_reset, int flags);

/* ------------------------------------------------------------------ hand-written policies -----
 * The reference's heuristic policies as stand-alone action generators over n_rows observation rows (one
 * row = one agent's observation), for callers that step an env one batch at a time; the
 * `*_rollout_heuristic` entry points above evaluate the same policies inside the rollout kernels.
 * madrl_ww_heuristic_actions       heuristics/waterworld.py:11-53: obs real [n_rows][obs_dim] in the 7K+2(+1)
 *                                  layout (:12-22) -> actions real [n_rows][2]; each row normalised by its own
 *                                  norm (the reference is called per agent).
 * madrl_pursuit_heuristic_actions  heuristics/pursuit.py:18-50: obs float [n_rows][obs_dim], flatten != 0: the
 *                                  (3R^2 [+1]) layout, else (R, R, 4); fallback int32 [n_rows] = the action taken
 *                                  when no evader is visible (`action_space.sample()`, drawn by the caller);
 *                                  floor_centre as in madrl_pursuit_rollout_heuristic; lut_dev: R*R bytes of
 *                                  device scratch (the per-cell action table, filled by the call). */
int madrl_ww_heuristic_actions(int fp64, size_t n_rows, int n_sensors, int obs_dim, const void* obs_dev,
                               void* actions_dev, void* stream);
int madrl_pursuit_heuristic_actions(size_t n_rows, int obs_range, int flatten, int obs_dim, int floor_centre,
                                    const float* obs_dev, const int32_t* fallback_dev, uint8_t* lut_dev,
                                    int32_t* actions_dev, void* stream);

/* ------------------------------------------------------------------ trajectory post-processing
 * (SURVEY.md 8f rows 2-3).  All tensors are device pointers, float32 unless noted, time-major. */
/* GAE advantages + discounted returns over paths segmented by `done`
 * (rllab/rllab/sampler/base.py:48-68).  rew, value, adv, ret: [T][E][A]; done: uint8 [T][E];
 * last_value [E][A] bootstraps the unfinished tail (NULL = 0, rllab's truncated-path convention). */
int madrl_gae_f32(int T, int E, int A, const float* rew_dev, const float* value_dev,
                  const uint8_t* done_dev, const float* last_value_dev, double discount,
                  double gae_lambda, float* adv_dev, float* ret_dev, void* stream);
/* ObservationBuffer frame stack (madrl_environments/__init__.py:143-196): obs [T][E][A][D] ->
 * out [T][E][A][D][B]; carry [E][A][D][B] is the buffer kept between calls; B <= 8. */
int madrl_frame_stack_f32(int T, int E, int A, int D, int B, const float* obs_dev,
                          const uint8_t* done_dev, float* carry_dev, float* out_dev, void* stream);
/* StandardizedEnv running mean / variance normalisation (madrl_environments/__init__.py:241-291),
 * in place over x [T][n] with float64 state mean/var [n]; center=1 for observations, 0 for rewards
 * (which are divided by the running std and multiplied by `scale`); enable=0 only applies `scale`. */
int madrl_standardize_f32(int T, size_t n, float* x_dev, double* mean_dev, double* var_dev,
                          double alpha, double eps, int center, double scale, int enable,
                          void* stream);
/* Same, for observations of an auto-reset rollout with the terminal observations kept on the side
 * (madrl_*_set_terminal_obs): at a step with done[t][e] set, the running estimate of env e's columns is
 * first updated with term[t] (which is standardised in place too) and then with x[t] -- the reference
 * order: step() standardises the terminal observation, reset() the first one of the next episode.
 * x, term [T][E][per_env]; done uint8 [T][E]; mean, var float64 [E][per_env]. */
int madrl_standardize_obs_terminal_f32(int T, int E, size_t per_env, float* x_dev, float* term_dev,
                                       const uint8_t* done_dev, double* mean_dev, double* var_dev,
                                       double alpha, double eps, void* stream);
/* DiagnosticsWrapper episode statistics (madrl_environments/__init__.py:314-369): rew [T][E][A],
 * done uint8 [T][E]; carry float64 [E][A+3] (zero-initialised by the caller, kept between calls);
 * outputs at the steps where an episode closes (ep_end[t][e] = 1; zeros elsewhere):
 * ep_reward [T][E][A], ep_disc [T][E] (discounted return of the agent-mean reward), ep_len [T][E]. */
int madrl_episode_stats_f32(int T, int E, int A, const float* rew_dev, const uint8_t* done_dev,
                            double discount, int max_traj_len, double* carry_dev, float* ep_reward_dev,
                            float* ep_disc_dev, int32_t* ep_len_dev, uint8_t* ep_end_dev, void* stream);

/* Whole-batch moments, deterministic (fixed grid, fixed combination order), float64 accumulation,
 * NumPy's two-pass population variance.  Series s0 = a, s1 = b, s2 = b - a (b may be NULL):
 * stats[0..2] = means, stats[3..5] = variances, stats[6..8] = minima (MADRL_MOMENTS_STATS doubles).
 * workspace: MADRL_MOMENTS_WS doubles of device scratch.  With a = baseline prediction and
 * b = returns this gives explained_variance_1d's terms (rllab/rllab/misc/special.py:51-59):
 * var(y) = stats[4], var(y - ypred) = stats[5], var(ypred) = stats[3]. */
#define MADRL_MOMENTS_STATS 9
#define MADRL_MOMENTS_WS 4096
int madrl_moments_f32(size_t n, const float* a_dev, const float* b_dev, double* stats_dev,
                      double* workspace_dev, void* stream);
/* center_advantages / shift_advantages_to_positive (rllab/rllab/algos/util.py:7-12, applied to the
 * concatenated advantages at rllab/rllab/sampler/base.py:82-86), in place over adv [n]:
 * center: (x - mean) / (std + 1e-8); positive: (x - min) + 1e-8 (after centring when both).
 * stats_dev receives the moments of the INPUT as in madrl_moments_f32 (series 0). */
int madrl_center_advantages_f32(size_t n, float* adv_dev, int center, int positive, double* stats_dev,
                                double* workspace_dev, void* stream);

/* Path packing: time-major rollout tensors -> rllab's per-(env, episode, agent) paths
 * (rllab/rllab/sampler/ma_sampler.py:52-100 `dec_rollout`, episodes cut at `done` as
 * rllab/sandbox/rocky/tf/envs/vec_env_executor.py:16-28 does).  Paths are ordered (env, episode, agent);
 * the rows of a path are consecutive in the packed arrays.
 *   madrl_paths_plan      done uint8 [T][E] -> seg_start / seg_len int32 [T][E] (start and length of the
 *                         episode containing step t), n_episodes int32 [E], ep_rec int32 [E][T][3] =
 *                         (start step, length, terminated) of the first n_episodes[e] episodes of env e;
 *                         path (e, episode j, agent a) occupies rows [e*T*A + s_j*A + a*L_j, +L_j)
 *   madrl_paths_pack_u32  src [T][E][A][D] of 4-byte words -> dst [E*T*A][D] in path order.  first != NULL:
 *                         rows are shifted by one step (row t = first[e][a] for t = 0, src[t-1] afterwards):
 *                         the observation each action was taken in, with first = the observations before
 *                         the rollout.  Per-env tensors (infos, dones) are packed with A = 1. */
int madrl_paths_plan(int T, int E, int A, const uint8_t* done_dev, int32_t* seg_start_dev, int32_t* seg_len_dev,
                     int32_t* n_episodes_dev, int32_t* ep_rec_dev, void* stream);
int madrl_paths_pack_u32(int T, int E, int A, int D, const void* src_dev, const void* first_dev,
                         const int32_t* seg_start_dev, const int32_t* seg_len_dev, void* dst_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MADRL_B200_H */
