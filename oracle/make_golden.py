"""Generate golden input/output vectors from the REAL reference (run in the build container).

    python -m oracle.make_golden

Drives the unmodified reference classes (via ``oracle/refshim.py``) with the counter-based
streams of ``oracle/philox.py`` injected where the reference draws random numbers, and writes
small ``.npz`` fixtures to ``tests/golden/``.  The reference tree does not exist on the GPU box;
these files (and this script, for provenance) are what travels.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle.philox import Stream  # noqa: E402
from oracle.refshim import load_reference  # noqa: E402

WW_CASES = {
    # name: (ctor kwargs, seed, env_id, T, action_std)
    "ww_c2": (dict(n_pursuers=5, n_evaders=5), 11, 3, 120, 0.5),
    "ww_dense": (dict(n_pursuers=5, n_evaders=5, n_coop=1, radius=0.04, sensor_range=0.3), 12, 0, 120, 1.0),
    "ww_c4": (dict(n_pursuers=20, n_evaders=50, n_poison=50), 13, 40000, 16, 0.5),
    "ww_c4_long": (dict(n_pursuers=20, n_evaders=50, n_poison=50), 15, 123456, 64, 0.7),
    "ww_global_nospeed": (dict(n_pursuers=3, n_evaders=4, n_poison=2, n_sensors=7, n_coop=1, radius=0.05,
                               reward_mech='global', speed_features=False, addid=False,
                               obstacle_loc=None), 14, 9, 150, 1.0),
}


class _StateLog(object):
    """Per-step snapshots of the REFERENCE env's internal state (entry t = state before step t,
    entry T = final state): the fp32 kernels are teacher-forced from these in the GPU tests."""

    def __init__(self, fn):
        self.fn, self.rows = fn, []

    def snap(self):
        self.rows.append(self.fn())

    def arrays(self):
        return {"st_" + k: np.array([r[k] for r in self.rows]) for k in self.rows[0]}


def _ww_state(env):
    objs = list(env._pursuers) + list(env._evaders) + list(env._poisons)
    return dict(pos=np.array([a.position for a in objs]), vel=np.array([a.velocity for a in objs]),
                obst=np.array(env.obstaclesx_No_2).reshape(2), t=env._timesteps,
                counter=env.np_random.counter)


def _hw_state(env):
    return dict(rx=np.array([a.position for a in env._rescuers]),
                rv=np.array([a.velocity for a in env._rescuers]),
                hx=np.array([a.position for a in env._hostages]),
                cx=np.array([a.position for a in env._criminals]),
                cv=np.array([a.velocity for a in env._criminals]),
                bomb=np.array(env.bomb_loc).reshape(2), key=np.array(env.key_loc).reshape(2),
                saved=np.array(env.curr_host_saved_mask, dtype=bool), gate=bool(env._gate_open),
                bombed=bool(env._bombed), t=env._timesteps, counter=env.np_random.counter)


def gen_waterworld(MAWaterWorld):
    for name, (kw, seed, env_id, T, std) in WW_CASES.items():
        env = MAWaterWorld(**kw)
        env.np_random = Stream(seed, env_id)
        Np = kw['n_pursuers']
        obs0 = np.array(env.reset())
        arng = np.random.RandomState(seed)
        actions = (arng.randn(T, Np, 2) * std).astype(np.float32).astype(np.float64)
        obs, rew, done, info = [], [], [], []
        st = _StateLog(lambda: _ww_state(env))
        for t in range(T):
            st.snap()
            o, r, d, i = env.step(actions[t])
            obs.append(np.array(o)); rew.append(np.array(r)); done.append(d)
            info.append([i['evcatches'], i['pocatches']])
        st.snap()
        cfg = {k: (None if v is None else v) for k, v in kw.items()}
        np.savez_compressed(
            os.path.join(GOLDEN, name + ".npz"), config=json.dumps(cfg), seed=seed, env_id=env_id,
            actions=actions, obs0=obs0, obs=np.array(obs), rew=np.array(rew),
            done=np.array(done), info=np.array(info, dtype=np.int32),
            final_px=np.array([p.position for p in env._pursuers]),
            final_ex=np.array([p.position for p in env._evaders]),
            final_ov=np.array([p.velocity for p in env._poisons]),
            counter=env.np_random.counter, **st.arrays())
        print(name, "catches", np.array(info).sum(0), "draws", env.np_random.counter)


PE_C3 = dict(n_evaders=30, n_pursuers=8, obs_range=7, surround=True, n_catch=2, flatten=True,
             reward_mech='local', catchr=0.1, term_pursuit=5.0, sample_maps=True, include_id=True)
PE_CASES = {
    # name: (map source, ctor kwargs, seed, env_id, T, resets every)
    "pe_c3": ("pool16", PE_C3, 21, 5, 160, 0),
    "pe_c3_global": ("pool16", dict(PE_C3, reward_mech='global', urgency_reward=-0.1), 22, 70000, 120, 50),
    "pe_ncatch": ("pool16", dict(PE_C3, surround=False, n_evaders=20, n_pursuers=12, obs_range=5), 23, 1, 160, 0),
    "pe_window": ("pool16", dict(PE_C3, constraint_window=0.5, n_evaders=6, n_pursuers=10, obs_range=9,
                                 include_id=False), 24, 2, 200, 70),
    "pe_small": ("small5", dict(n_evaders=2, n_pursuers=4, obs_range=3, surround=True, reward_mech='local',
                                catchr=0.1, sample_maps=False), 25, 3, 1500, 0),
    "pe_crowd": ("small5", dict(n_evaders=4, n_pursuers=10, obs_range=3, surround=True, reward_mech='local',
                                catchr=0.1, term_pursuit=5.0, sample_maps=False), 27, 6, 600, 0),
    "pe_random_opp": ("small5", dict(n_evaders=5, n_pursuers=6, obs_range=3, surround=False, n_catch=1,
                                     reward_mech='local', catchr=0.1, random_opponents=True, max_opponents=6),
                      28, 7, 400, 0),
    "pe_even_range": ("small5", dict(n_evaders=3, n_pursuers=3, obs_range=4, surround=False, n_catch=1,
                                     reward_mech='global'), 26, 4, 200, 0),
}


def small_map():
    m = np.zeros((1, 5, 5), dtype=np.int32)
    m[0, 2, 2] = -1
    return m


def gen_pursuit():
    from oracle.refshim import make_reference_pursuit, REFERENCE_ROOT
    pool16 = np.load(os.path.join(REFERENCE_ROOT, "maps", "map_pool16.npy"))
    # the map pool is an INPUT fixture of BASELINE.json configs 1 and 3 (reference file
    # maps/map_pool16.npy, int32 (10,16,16), -1 = building); kept at the same relative path
    os.makedirs(os.path.join(ROOT, "maps"), exist_ok=True)
    np.save(os.path.join(ROOT, "maps", "map_pool16.npy"), pool16)
    for name, (msrc, kw, seed, env_id, T, every) in PE_CASES.items():
        maps = pool16 if msrc == "pool16" else small_map()
        stream = Stream(seed, env_id)
        env = make_reference_pursuit(maps, stream, **kw)
        Np = kw['n_pursuers']
        arng = np.random.RandomState(seed)
        actions = arng.randint(0, 5, size=(T, Np)).astype(np.int32)
        obs0 = np.array(env.reset())
        obs, rew, done, removed, reset_at, reset_obs = [], [], [], [], [], []
        for t in range(T):
            o, r, d, i = env.step(actions[t])
            obs.append(np.array(o)); rew.append(np.asarray(r, dtype=np.float64)); done.append(d)
            removed.append(i['removed'])
            if d or (every and t % every == every - 1):
                reset_at.append(t)
                reset_obs.append(np.array(env.reset()))
        np.savez_compressed(
            os.path.join(GOLDEN, name + ".npz"), config=json.dumps(kw), maps=msrc, seed=seed,
            env_id=env_id, actions=actions, obs0=obs0, obs=np.array(obs), rew=np.array(rew),
            done=np.array(done), removed=np.array(removed, dtype=np.int32),
            reset_at=np.array(reset_at, dtype=np.int32),
            reset_obs=np.array(reset_obs) if reset_obs else np.zeros((0,) + obs0.shape),
            counter=stream.counter)
        print(name, "removed", int(np.sum(removed)), "dones", int(np.sum(done)), "resets", len(reset_at),
              "draws", stream.counter)


HW_CASES = {
    # name: (positional args, kwargs, seed, env_id, T, action_std)
    "hw_c5": ((10, 16, 16, 4, 2), {}, 31, 8, 120, 0.5),
    "hw_c5_local": ((10, 16, 16, 4, 2), dict(reward_mech='local'), 32, 9000, 150, 2.0),
    "hw_dense": ((3, 10, 5, 1, 2), dict(radius=0.05, sensor_range=0.35, key_radius=0.06, reward_mech='local',
                                        addid=False), 33, 1, 400, 3.0),
    "hw_k12": ((4, 6, 8, 2, 1), dict(radius=0.04, n_sensors=12, key_radius=0.05, bomb_radius=0.02,
                                     key_loc=[[0.93, 0.97]]), 34, 2, 400, 3.0),
}


def gen_hostage(ContinuousHostageWorld):
    for name, (args, kw, seed, env_id, T, std) in HW_CASES.items():
        kw2 = dict(kw)
        if 'key_loc' in kw2:
            kw2['key_loc'] = np.array(kw2['key_loc'])
        env = ContinuousHostageWorld(*args, **kw2)
        env.np_random = Stream(seed, env_id)
        Nr = args[0]
        obs0 = np.array(env.reset())
        arng = np.random.RandomState(seed)
        actions = (arng.randn(T, Nr, 2) * std).astype(np.float32).astype(np.float64)
        obs, rew, done, info, reset_at, reset_obs = [], [], [], [], [], []
        st = _StateLog(lambda: _hw_state(env))
        for t in range(T):
            st.snap()
            o, r, d, i = env.step(actions[t])
            obs.append(np.array(o)); rew.append(np.array(r)); done.append(d)
            info.append([i['ho_saved'], i['cr_encs']])
            if d:
                reset_at.append(t)
                reset_obs.append(np.array(env.reset()))
        st.snap()
        np.savez_compressed(
            os.path.join(GOLDEN, name + ".npz"), args=np.array(args), config=json.dumps(kw), seed=seed,
            env_id=env_id, actions=actions, obs0=obs0, obs=np.array(obs), rew=np.array(rew),
            done=np.array(done), info=np.array(info, dtype=np.int32),
            reset_at=np.array(reset_at, dtype=np.int32),
            reset_obs=np.array(reset_obs) if reset_obs else np.zeros((0,) + obs0.shape),
            counter=env.np_random.counter, **st.arrays())
        print(name, "saved/encs", np.array(info).sum(0), "dones", int(np.sum(done)), "draws", env.np_random.counter)


CL_WW_CASES = {
    # name: (ctor kwargs, seed, env_id, T)
    "cl_ww_c2": (dict(n_pursuers=5, n_evaders=5), 41, 17, 150),
    # seed chosen so that no decision normalises a rounding residue: with seed 42 two pursuers sense equal distances
    # in opposite directions at step 44, the weighted sum is 2e-18 instead of 0 and the reference policy turns that
    # noise into a unit action -- reproducible only with bit-identical observations (min_norm is recorded below)
    "cl_ww_dense": (dict(n_pursuers=4, n_evaders=6, n_poison=6, n_coop=1, radius=0.04, sensor_range=0.3), 45, 2, 150),
}
CL_PE_CASES = {
    # name: (map source, ctor kwargs (flatten=False: the policy reads the (R, R, 4) layout), seed, env_id, T, py2 division)
    "cl_pe_conv_py2": ("pool16", dict(PE_C3, flatten=False, n_evaders=12, obs_range=7), 43, 11, 120, True),
    "cl_pe_conv_py3": ("pool16", dict(PE_C3, flatten=False, n_evaders=12, obs_range=7), 43, 11, 120, False),
    "cl_pe_sparse_py2": ("pool16", dict(PE_C3, flatten=False, n_evaders=3, n_pursuers=6, obs_range=5, surround=False,
                                        n_catch=1), 44, 12, 150, True),
}


def gen_closed_loop(MAWaterWorld):
    """CLOSED LOOP of the real reference: the reference env stepped with the actions of the reference's own
    hand-written policy (heuristics/waterworld.py, heuristics/pursuit.py), called per agent as the reference's
    Visualizer does.  The engine's in-kernel policy rollouts (madrl_*_rollout_heuristic) must reproduce these
    trajectories.  Pursuit: `action_space.sample()` is the injected policy stream of oracle/heuristics_oracle.py;
    py2 = line 23's `xs / 2, ys / 2` spelt `//` (what it means in the reference's Python 2)."""
    import types
    from oracle.heuristics_oracle import policy_draw
    from oracle.refshim import REFERENCE_ROOT, load_reference_heuristics, make_reference_pursuit
    WPol, PPol, psrc = load_reference_heuristics()
    for name, (kw, seed, env_id, T) in CL_WW_CASES.items():
        env = MAWaterWorld(**kw)
        env.np_random = Stream(seed, env_id)
        pol = WPol(env.agents[0].observation_space, env.agents[0].action_space)
        o = env.reset()
        obs0 = np.array(o)
        acts, obs, rew, info, min_norm = [], [], [], [], np.inf
        from oracle.heuristics_oracle import waterworld_action
        for t in range(T):
            a = np.array([pol.sample_actions(np.asarray(oi)[None])[0][0] for oi in o])
            norms = [waterworld_action(np.asarray(oi), True)[1] for oi in o]
            min_norm = min([min_norm] + [x for x in norms if x > 0])
            o, r, d, i = env.step(a)
            acts.append(a); obs.append(np.array(o)); rew.append(np.array(r)); info.append([i['evcatches'], i['pocatches']])
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), config=json.dumps(kw), seed=seed, env_id=env_id,
                            obs0=obs0, actions=np.array(acts), obs=np.array(obs), rew=np.array(rew),
                            info=np.array(info, dtype=np.int32), counter=env.np_random.counter, min_norm=min_norm)
        assert min_norm > 1e-6, "ill-conditioned closed loop (a rounding residue is normalised): pick another seed"
        print(name, "catches", np.array(info).sum(0), "moving", float(np.mean(np.linalg.norm(acts, axis=-1) > 0)),
              "min un-normalised norm", min_norm)
    pool16 = np.load(os.path.join(REFERENCE_ROOT, "maps", "map_pool16.npy"))
    ported = psrc.replace("x, y = xs / 2, ys / 2", "x, y = xs // 2, ys // 2")
    assert ported != psrc
    mod2 = types.ModuleType("_ref_heuristics_pursuit_py2")
    exec(compile(ported.split("if __name__")[0], "heuristics/pursuit.py[py2 division]", "exec"), mod2.__dict__)

    class _Space(object):
        def sample(self):
            return self.draw()

    for name, (msrc, kw, seed, env_id, T, py2) in CL_PE_CASES.items():
        maps = pool16 if msrc == "pool16" else small_map()
        stream = Stream(seed, env_id)
        env = make_reference_pursuit(maps, stream, **kw)
        space = _Space()
        pol = (mod2.PursuitHeuristicPolicy if py2 else PPol)(None, space)
        o = env.reset()
        obs0 = np.array(o)
        acts, obs, rew, done, removed, sampled = [], [], [], [], [], 0
        for t in range(T):
            a = []
            for q, oq in enumerate(o):
                hit = []
                space.draw = lambda q=q: hit.append(1) or policy_draw(seed, env_id, stream.counter, q)
                a.append(int(pol.sample_actions(np.asarray(oq))[0]))
                sampled += len(hit)
            o, r, d, i = env.step(a)
            acts.append(a); obs.append(np.array(o)); rew.append(np.asarray(r, dtype=np.float64)); done.append(d)
            removed.append(i['removed'])
            if d:
                break
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), config=json.dumps(kw), maps=msrc, seed=seed,
                            env_id=env_id, py2=py2, obs0=obs0, actions=np.array(acts, dtype=np.int32),
                            obs=np.array(obs), rew=np.array(rew), done=np.array(done),
                            removed=np.array(removed, dtype=np.int32), counter=stream.counter)
        print(name, "steps", len(acts), "removed", int(np.sum(removed)), "sampled actions", sampled)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    MAWaterWorld, PursuitEvade, ContinuousHostageWorld = load_reference()
    gen_waterworld(MAWaterWorld)
    gen_pursuit()
    gen_hostage(ContinuousHostageWorld)
    gen_closed_loop(MAWaterWorld)


if __name__ == "__main__":
    main()
