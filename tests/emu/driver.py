"""TEST INFRASTRUCTURE: ctypes drivers for the warp-emulator build of the kernels
(tests/emu/build_emu.py).  The emulator library exports the product's C ABI; "device" pointers are
host pointers, so the entry points are called directly with NumPy arrays.  Argument names and
defaults mirror madrl_b200.{waterworld,pursuit,hostage}.Batched* so the tests read like the GPU ones.
"""
import ctypes as C

import numpy as np

from madrl_b200 import _lib as L

from . import build_emu

_libs = {}


def load(defines=()):
    key = tuple(sorted(defines))
    if key not in _libs:
        lib = C.CDLL(build_emu.build(list(key)))
        _declare(lib)
        _libs[key] = lib
    return _libs[key]


def races(reset=True):
    """Shared-memory hazards (see smem_access in tests/emu/include/cuda_runtime.h) reported by every
    loaded emulator library since the last call."""
    n = 0
    for lib in _libs.values():
        n += int(lib.madrl_emu_race_count())
        if reset:
            lib.madrl_emu_race_reset()
    return n


def _declare(lib):
    vp, i32, u64 = C.c_void_p, C.c_int, C.c_uint64
    lib.madrl_last_error.restype = C.c_char_p
    lib.madrl_emu_race_count.restype = C.c_long
    for fam, cfg, lay in (("ww", L.WWConfig, L.WWLayout), ("hostage", L.HWConfig, L.HWLayout),
                          ("pursuit", L.PEConfig, L.PELayout)):
        g = lambda n: getattr(lib, "madrl_%s_%s" % (fam, n))   # noqa: E731
        g("state_layout").argtypes = [C.POINTER(cfg), C.POINTER(lay)]
        g("create").argtypes = [C.POINTER(cfg), vp, vp, C.POINTER(vp)] if fam == "pursuit" else \
            [C.POINTER(cfg), vp, C.POINTER(vp)]
        g("destroy").argtypes = [vp]
        g("seed").argtypes = [vp, u64, vp]
        g("set_launch").argtypes = [vp, i32, i32]
        g("reset").argtypes = [vp, vp, vp, vp]
        g("rollout").argtypes = [vp, i32, vp, vp, vp, vp, vp, i32, vp]
        g("rollout_host").argtypes = [vp, i32, vp, vp, vp, vp, vp, i32]
        g("reset_host").argtypes = [vp, vp, vp]
    lib.madrl_ww_rollout_heuristic.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, i32, vp]
    lib.madrl_pursuit_rollout_heuristic.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, vp]


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


GUARD = 1 << 16   # bytes of canary on either side of every output buffer


class Guarded(object):
    """An output array embedded in a canary-filled allocation: out-of-row writes of a kernel show up
    as a failed check instead of silent corruption."""

    def __init__(self, shape, dtype, fill):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.raw = np.full(n + 2 * GUARD, 0xA5, np.uint8)
        self.arr = self.raw[GUARD:GUARD + n].view(dtype).reshape(shape)
        self.arr[...] = fill

    def check(self, what):
        assert (self.raw[:GUARD] == 0xA5).all() and (self.raw[-GUARD:] == 0xA5).all(), \
            "kernel wrote outside the %s buffer" % what
        return self.arr


class _Engine(object):
    fam = None

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("emulated madrl_b200 error %d: %s" % (rc, self.lib.madrl_last_error().decode()))

    def _f(self, name):
        return getattr(self.lib, "madrl_%s_%s" % (self.fam, name))

    def _finish(self, create_args):
        self.layout = self.LayoutT()
        self._check(self._f("state_layout")(C.byref(self.cfg), C.byref(self.layout)))
        self.obs_dim = int(self.layout.obs_dim)
        self.blob = np.zeros(int(self.layout.total_bytes), np.uint8)
        h = C.c_void_p()
        self._check(self._f("create")(C.byref(self.cfg), *create_args, _p(self.blob), C.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._f("destroy")(h)

    def view(self, off, dtype, shape):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        return self.blob[off:off + n].view(dtype).reshape(shape)

    def set_launch(self, warps_per_block=0, blocks_per_sm=0):
        self._check(self._f("set_launch")(self._h, warps_per_block, blocks_per_sm))

    def seed(self, seed):
        self._check(self._f("seed")(self._h, int(seed), None))

    def reset(self, mask=None):
        obs = Guarded((self.n_envs, self.n_agents, self.obs_dim), self.obs_dtype, 0)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self._check(self._f("reset")(self._h, _p(m), _p(obs.arr), None))
        return obs.check("obs")

    def rollout(self, actions, auto_reset=True, host=False):
        actions = np.ascontiguousarray(actions, self.act_dtype)
        T = actions.shape[0]
        E, A = self.n_envs, self.n_agents
        bufs = [Guarded((T, E, A, self.obs_dim), self.obs_dtype, np.nan), Guarded((T, E, A), self.obs_dtype, np.nan),
                Guarded((T, E), np.uint8, 255), Guarded((T, E) + self.info_shape, np.int32, -1)]
        ptrs = [_p(b.arr) for b in bufs]
        if host:
            self._check(self._f("rollout_host")(self._h, T, _p(actions), *ptrs, int(auto_reset)))
        else:
            self._check(self._f("rollout")(self._h, T, _p(actions), *ptrs, int(auto_reset), None))
        return tuple(b.check(n) for b, n in zip(bufs, ("obs", "rew", "done", "info")))


    def rollout_heuristic(self, T, obs0, auto_reset=True, **kw):
        """Closed-loop rollout with the in-kernel heuristic policy -> (actions, obs, rew, done, info)."""
        E, A = self.n_envs, self.n_agents
        obs0 = np.ascontiguousarray(obs0, self.obs_dtype)
        assert obs0.shape == (E, A, self.obs_dim)
        bufs = [Guarded((T, E, A) + self.act_shape, self.act_dtype, -7),
                Guarded((T, E, A, self.obs_dim), self.obs_dtype, np.nan), Guarded((T, E, A), self.obs_dtype, np.nan),
                Guarded((T, E), np.uint8, 255), Guarded((T, E) + self.info_shape, np.int32, -1)]
        ptrs = [_p(b.arr) for b in bufs]
        extra = [int(bool(kw.get("py2_division", True)))] if self.fam == "pursuit" else []
        self._check(self._f("rollout_heuristic")(self._h, T, _p(obs0), *ptrs, int(auto_reset), *extra, None))
        return tuple(b.check(n) for b, n in zip(bufs, ("actions", "obs", "rew", "done", "info")))


def ww_heuristic_actions(obs, n_sensors, defines=()):
    """madrl_ww_heuristic_actions through the emulator: obs [n, D] float32 / float64 -> actions [n, 2]."""
    lib = load(defines)
    lib.madrl_ww_heuristic_actions.argtypes = [C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    obs = np.ascontiguousarray(obs)
    out = Guarded((obs.shape[0], 2), obs.dtype, np.nan)
    rc = lib.madrl_ww_heuristic_actions(int(obs.dtype == np.float64), obs.shape[0], n_sensors, obs.shape[1], _p(obs),
                                        _p(out.arr), None)
    assert rc == 0, lib.madrl_last_error().decode()
    return out.check("actions")


def pursuit_heuristic_actions(obs, obs_range, flatten, fallback, py2_division=True, defines=()):
    """madrl_pursuit_heuristic_actions through the emulator: obs [n, D] float32 -> actions [n] int32."""
    lib = load(defines)
    lib.madrl_pursuit_heuristic_actions.argtypes = [C.c_size_t] + [C.c_int] * 4 + [C.c_void_p] * 5
    obs = np.ascontiguousarray(obs, np.float32)
    fb = np.ascontiguousarray(fallback, np.int32)
    lut = np.zeros(128, np.uint8)
    out = Guarded((obs.shape[0],), np.int32, -9)
    rc = lib.madrl_pursuit_heuristic_actions(obs.shape[0], obs_range, int(flatten), obs.shape[1], int(py2_division),
                                             _p(obs), _p(fb), _p(lut), _p(out.arr), None)
    assert rc == 0, lib.madrl_last_error().decode()
    return out.check("actions")


class EmuWaterworld(_Engine):
    fam, LayoutT, info_shape, act_shape = "ww", L.WWLayout, (2,), (2,)

    def __init__(self, n_envs, n_pursuers, n_evaders, n_coop=2, n_poison=10, radius=0.015,
                 obstacle_radius=0.2, obstacle_loc=np.array([0.5, 0.5]), ev_speed=0.01,
                 poison_speed=0.01, n_sensors=30, sensor_range=0.2, action_scale=0.01,
                 poison_reward=-1., food_reward=1., encounter_reward=.05, control_penalty=-.5,
                 reward_mech='local', addid=True, speed_features=True, seed=0, env_id_base=0,
                 max_path_length=0, fp64=True, defines=()):
        self.lib = load(defines)
        self.n_envs, self.n_agents, self.n_pursuers, self.n_evaders = n_envs, n_pursuers, n_pursuers, n_evaders
        self.obs_dtype = self.act_dtype = np.float64 if fp64 else np.float32
        rand = obstacle_loc is None
        ox, oy = (0.0, 0.0) if rand else (float(obstacle_loc[0]), float(obstacle_loc[1]))
        self.cfg = L.WWConfig(
            n_envs=n_envs, env_id_base=env_id_base, n_pursuers=n_pursuers, n_evaders=n_evaders,
            n_poison=n_poison, n_sensors=n_sensors, n_coop=n_coop,
            reward_global=int(reward_mech == 'global'), addid=int(bool(addid)),
            speed_features=int(bool(speed_features)), random_obstacle=int(rand), timestep_limit=1000,
            max_path_length=int(max_path_length or 0), fp64=int(fp64), radius=radius,
            obstacle_radius=obstacle_radius, obstacle_x=ox, obstacle_y=oy, ev_speed=ev_speed,
            poison_speed=poison_speed, sensor_range=sensor_range, action_scale=action_scale,
            poison_reward=poison_reward, food_reward=food_reward, encounter_reward=encounter_reward,
            control_penalty=control_penalty, seed=int(seed))
        self._finish(())

    def state(self, e):
        Ly, N = self.layout, int(self.layout.n_obj)
        objs = self.view(Ly.objs, self.obs_dtype, (self.n_envs, 4, N))[e].astype(np.float64)
        X, V = objs[0:2].T, objs[2:4].T
        Np, Ne = self.n_pursuers, self.n_evaders
        obst = self.view(Ly.obst, self.obs_dtype, (self.n_envs, 2))[e].astype(np.float64)[None]
        return dict(px=X[:Np], pv=V[:Np], ex=X[Np:Np + Ne], ev=V[Np:Np + Ne], ox=X[Np + Ne:], ov=V[Np + Ne:],
                    obst=obst, t=int(self.view(Ly.timestep, np.int32, (self.n_envs,))[e]),
                    counter=int(self.view(Ly.rng_counter, np.int64, (self.n_envs,))[e]))


class EmuHostage(_Engine):
    fam, LayoutT, info_shape, act_shape = "hostage", L.HWLayout, (2,), (2,)

    def __init__(self, n_envs, n_good, n_hostages, n_bad, n_coop_save, n_coop_avoid, radius=0.015,
                 key_loc=None, bad_speed=0.01, n_sensors=30, sensor_range=0.2, action_scale=0.01,
                 save_reward=5., hit_reward=-1., encounter_reward=0.01, not_saved_reward=-3,
                 bomb_reward=-5., bomb_radius=0.05, key_radius=0.0075, control_penalty=-.1,
                 reward_mech='global', addid=True, seed=0, env_id_base=0, max_path_length=0, fp64=True,
                 defines=()):
        self.lib = load(defines)
        self.n_envs, self.n_agents, self.n_hostages = n_envs, n_good, n_hostages
        self.obs_dtype = self.act_dtype = np.float64 if fp64 else np.float32
        rand_key = key_loc is None
        kx, ky = (0.0, 0.0) if rand_key else [float(v) for v in np.asarray(key_loc).reshape(-1)[:2]]
        self.cfg = L.HWConfig(
            n_envs=n_envs, env_id_base=env_id_base, n_good=n_good, n_hostages=n_hostages, n_bad=n_bad,
            n_coop_save=n_coop_save, n_coop_avoid=n_coop_avoid, n_sensors=n_sensors,
            reward_global=int(reward_mech == 'global'), addid=int(bool(addid)), random_key=int(rand_key),
            timestep_limit=1000, max_path_length=int(max_path_length or 0), fp64=int(fp64), radius=radius,
            key_x=kx, key_y=ky, bad_speed=bad_speed, sensor_range=float(sensor_range),
            action_scale=action_scale, save_reward=save_reward, hit_reward=hit_reward,
            encounter_reward=encounter_reward, not_saved_reward=float(not_saved_reward),
            bomb_reward=bomb_reward, bomb_radius=bomb_radius, key_radius=key_radius,
            control_penalty=control_penalty, seed=int(seed))
        self._finish(())

    def state(self, e):
        Ly, E = self.layout, self.n_envs
        return dict(counter=int(self.view(Ly.rng_counter, np.int64, (E,))[e]),
                    t=int(self.view(Ly.timestep, np.int32, (E,))[e]),
                    flags=int(self.view(Ly.flags, np.int32, (E,))[e]),
                    saved=self.view(Ly.saved, np.uint8, (E, self.n_hostages))[e].astype(bool))


class EmuPursuit(_Engine):
    fam, LayoutT, info_shape, act_shape = "pursuit", L.PELayout, (), ()
    obs_dtype, act_dtype = np.float32, np.int32

    def __init__(self, n_envs, map_pool, n_evaders=1, n_pursuers=1, obs_range=3, flatten=True,
                 layer_norm=10, n_catch=2, catchr=0.01, term_pursuit=5.0, urgency_reward=0.0,
                 include_id=True, surround=True, constraint_window=1.0, sample_maps=False,
                 reward_mech='global', random_opponents=False, max_opponents=10, seed=0, env_id_base=0,
                 max_path_length=0, defines=()):
        self.lib = load(defines)
        mp = np.ascontiguousarray(np.asarray(map_pool), dtype=np.int32)
        if mp.ndim == 2:
            mp = mp[None]
        self.map_pool = mp
        self.n_envs, self.n_agents, self.n_pursuers, self.n_evaders = n_envs, n_pursuers, n_pursuers, n_evaders
        self.cfg = L.PEConfig(
            n_envs=n_envs, env_id_base=env_id_base, n_pursuers=n_pursuers, n_evaders=n_evaders,
            xs=mp.shape[1], ys=mp.shape[2], n_maps=mp.shape[0], obs_range=obs_range,
            flatten=int(bool(flatten)), n_catch=n_catch, surround=int(bool(surround)),
            reward_global=int(reward_mech == 'global'), include_id=int(bool(include_id)),
            sample_maps=int(bool(sample_maps)), max_path_length=int(max_path_length or 0),
            max_opponents=int(max_opponents) if random_opponents else 0,
            layer_norm=float(layer_norm), catchr=float(catchr), term_pursuit=float(term_pursuit),
            urgency_reward=float(urgency_reward), constraint_window=float(constraint_window), seed=int(seed))
        self._finish((_p(mp),))

    def state(self, e):
        Ly, E, Np = self.layout, self.n_envs, self.n_pursuers
        A = int(Ly.n_agents)
        pos = self.view(Ly.pos, np.uint8, (E, 2, A))[e]
        return dict(pursuers=pos[:, :Np].T.astype(int), evaders=pos[:, Np:].T.astype(int),
                    gone=int(self.view(Ly.gone, np.int64, (E,))[e]),
                    map_id=int(self.view(Ly.map_id, np.int32, (E,))[e]),
                    counter=int(self.view(Ly.rng_counter, np.int64, (E,))[e]))


# ---------------------------------------------------------------------------------------------------
# Test double for the DROP-IN classes (madrl_b200.MAWaterWorld / PursuitEvade / ContinuousHostageWorld):
# those classes talk to a `Batched*` engine (reset / step / seed / set_params returning torch tensors).
# `emulated_dropins()` swaps the engine classes for facades over the emulator engines above, so the
# drop-in code itself -- list-of-arrays returns, float64 conversion, info dicts, spaces, pickling,
# the vec_env_executor hook -- runs unchanged on a CPU box, against the reference's own callers
# (tests/test_reference_callers.py).  On a GPU box the same classes run on the real library.
# ---------------------------------------------------------------------------------------------------
class _TorchFacade(object):
    def __init__(self, emu, info_keys):
        import torch
        self._t, self.emu, self._keys = torch, emu, info_keys
        self.n_envs, self.obs_dim = emu.n_envs, emu.obs_dim
        self.device = torch.device("cpu")
        self.dtype = torch.float64 if emu.obs_dtype == np.float64 else torch.float32

    def seed(self, seed=None):
        self.emu.seed(0 if seed is None else int(seed))
        return [seed]

    def reset(self, mask=None, out=None):
        return self._t.from_numpy(np.array(self.emu.reset(None if mask is None else np.asarray(mask))))

    def step(self, actions, auto_reset=False):
        a = np.asarray(actions).reshape((1, self.n_envs) + self._act_tail)
        obs, rew, done, info = [self._t.from_numpy(np.array(x)) for x in self.emu.rollout(a, auto_reset=auto_reset)]
        if len(self._keys) == 1:
            d = {self._keys[0]: info[0]}
        else:
            d = {k: info[0, :, i] for i, k in enumerate(self._keys)}
        return obs[0], rew[0], done[0], d


def _ww_facade(n_envs, *args, device=None, dtype=None, **kw):
    import torch
    f = _TorchFacade(EmuWaterworld(n_envs, *args, fp64=(dtype == torch.float64), **kw), ("evcatches", "pocatches"))
    f._act_tail = (f.emu.n_pursuers, 2)
    return f


def _hw_facade(n_envs, *args, device=None, dtype=None, **kw):
    import torch
    f = _TorchFacade(EmuHostage(n_envs, *args, fp64=(dtype == torch.float64), **kw), ("ho_saved", "cr_encs"))
    f._act_tail = (f.emu.n_agents, 2)
    return f


def _pe_facade(n_envs, map_pool, device=None, **kw):
    f = _TorchFacade(EmuPursuit(n_envs, map_pool, **kw), ("removed",))
    f._act_tail = (f.emu.n_pursuers,)

    def set_params(catchr, constraint_window):
        lib = f.emu.lib
        lib.madrl_pursuit_set_params.argtypes = [C.c_void_p, C.c_double, C.c_double]
        f.emu._check(lib.madrl_pursuit_set_params(f.emu._h, float(catchr), float(constraint_window)))
    f.set_params = set_params
    import torch
    type(f).state = property(lambda self: {'gone': torch.tensor([self.emu.state(e)['gone'] for e in range(self.n_envs)])})
    return f


class emulated_dropins(object):
    """Context manager: madrl_b200's drop-in env classes run on the emulator engines."""

    def __enter__(self):
        import madrl_b200.hostage as H
        import madrl_b200.pursuit as P
        import madrl_b200.waterworld as W
        self._saved = (W.BatchedMAWaterWorld, P.BatchedPursuitEvade, H.BatchedHostageWorld)
        W.BatchedMAWaterWorld, P.BatchedPursuitEvade, H.BatchedHostageWorld = _ww_facade, _pe_facade, _hw_facade
        return self

    def __exit__(self, *exc):
        import madrl_b200.hostage as H
        import madrl_b200.pursuit as P
        import madrl_b200.waterworld as W
        W.BatchedMAWaterWorld, P.BatchedPursuitEvade, H.BatchedHostageWorld = self._saved
        return False
