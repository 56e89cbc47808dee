"""The drop-in boundary exercised by the REFERENCE'S OWN callers (needs /root/reference; CPU).

INTEGRATION.md claims that the existing runners / samplers run unchanged on the drop-in classes.
Here the real caller code is imported (``oracle/refshim.py``: no reference file is copied or edited)
and driven twice -- once over the reference env, once over ``madrl_b200``'s drop-in class -- on the
same injected Philox stream and the same scripted policy; everything the callers produce must agree:

  * the env wrappers ``ObservationBuffer`` / ``StandardizedEnv`` / ``DiagnosticsWrapper``
    (madrl_environments/__init__.py:143-389),
  * rltools ``decrollout`` (rltools/rltools/samplers/__init__.py:148-190),
  * ``RLLabEnv`` + rllab ``dec_rollout`` (rllabwrapper/__init__.py:29-90, rllab/rllab/sampler/ma_sampler.py:52-100),
  * rllab ``VecEnvExecutor`` (rllab/sandbox/rocky/tf/envs/vec_env_executor.py:6-48) vs the
    ``env.vec_env_executor(n_envs, max_path_length)`` hook.

The drop-in classes run on the emulator build of the kernels here (tests/emu ``emulated_dropins``:
only the engine object is swapped, the class code is the product's); under ``-m gpu`` the same
classes are covered on the real library by tests/test_*_gpu.py::test_dropin_env_surface.
"""
import numpy as np
import pytest
import torch

from oracle.philox import Stream

pytestmark = pytest.mark.reference

WW_KW = dict(n_pursuers=3, n_evaders=4, n_poison=3, n_sensors=8, n_coop=1, radius=0.05, sensor_range=0.3)
HW_ARGS, HW_KW = (3, 5, 4, 1, 1), dict(radius=0.05, n_sensors=8, sensor_range=0.3, key_radius=0.06)
SEED = 31


@pytest.fixture()
def world():
    """Reference classes + drop-in classes with gym's space classes shared (as on a machine where gym
    is installed: madrl_b200/spaces.py then re-exports gym.spaces.Box / Discrete)."""
    from oracle.refshim import install, load_reference
    install()
    import gym.spaces as gs
    import madrl_b200
    from madrl_b200 import hostage as H, pursuit as P, spaces as S, waterworld as W
    from emu.driver import emulated_dropins
    saved = [(m, m.Box) for m in (S, W, P, H)] + [(m, None) for m in ()]
    saved_d = [(m, m.Discrete) for m in (S, P)]
    for m, _ in saved:
        m.Box = gs.Box
    for m, _ in saved_d:
        m.Discrete = gs.Discrete
    ref = load_reference()
    with emulated_dropins():
        yield dict(ref=ref, ours=(madrl_b200.MAWaterWorld, madrl_b200.PursuitEvade, madrl_b200.ContinuousHostageWorld))
    for m, b in saved:
        m.Box = b
    for m, d in saved_d:
        m.Discrete = d


def make_pair(world, family, env_id=0):
    """(reference env, drop-in env) of one family on the same stream key (SEED, env_id)."""
    RefWW, _, RefHW = world["ref"]
    OurWW, _, OurHW = world["ours"]
    if family == "ww":
        ref = RefWW(**WW_KW)
        ref.np_random = Stream(SEED, env_id)
        return ref, OurWW(seed=SEED, env_id=env_id, dtype=torch.float64, **WW_KW)
    ref = RefHW(*HW_ARGS, **HW_KW)
    ref.np_random = Stream(SEED, env_id)
    return ref, OurHW(*HW_ARGS, seed=SEED, env_id=env_id, dtype=torch.float64, **HW_KW)


def make_pursuit_pair(world):
    from oracle.refshim import make_reference_pursuit
    maps = np.zeros((1, 6, 6), dtype=np.int32)
    maps[0, 2, 3] = -1
    kw = dict(n_evaders=3, n_pursuers=4, obs_range=3, surround=False, n_catch=1, reward_mech='local', catchr=0.1,
              sample_maps=False)
    ref = make_reference_pursuit(maps, Stream(SEED, 5), **kw)
    return ref, world["ours"][1](maps, seed=SEED, env_id=5, **kw)


def same(a, b, tol=1e-9):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and (a.size == 0 or float(np.abs(a - b).max()) <= tol)


# ------------------------------------------------------------------------------------ wrappers
@pytest.mark.parametrize("family", ["ww", "hw"])
def test_reference_wrappers_around_dropin_equal_wrappers_around_reference(world, family):
    import madrl_environments as me
    ref, ours = make_pair(world, family)
    n_agents = len(ref.agents)
    assert len(ours.agents) == n_agents and ours.reward_mech == ref.reward_mech
    for a, b in zip(ref.agents, ours.agents):      # what the wrappers read at construction
        assert a.observation_space.shape == b.observation_space.shape and a.action_space.shape == b.action_space.shape
        assert np.array_equal(a.observation_space.low, b.observation_space.low)

    def stack(env):
        # ObservationBuffer.__init__ assigns to its own read-only `reward_mech` property
        # (madrl_environments/__init__.py:151 vs :171) and raises for ANY env, the reference's included;
        # the runners never enable it (buffer_size 1).  Its step/reset logic is exercised by building the
        # object the way its __init__ would have.
        with pytest.raises(AttributeError):
            me.ObservationBuffer(env, 3)
        ob = me.ObservationBuffer.__new__(me.ObservationBuffer)
        ob._unwrapped, ob._buffer_size = env, 3
        ob._buffer = [np.zeros(tuple(ag.observation_space.shape) + (3,)) for ag in env.agents]
        std = me.StandardizedEnv(env, scale_reward=0.5, enable_obsnorm=True, enable_rewnorm=True, obs_alpha=0.05,
                                 rew_alpha=0.02)
        return ob, std, me.DiagnosticsWrapper(std, discount=0.9, max_traj_len=20, log_interval=10 ** 9)

    # two independent copies per side: ObservationBuffer and the StandardizedEnv/Diagnostics chain both step the env
    ref2, ours2 = make_pair(world, family)
    ob_r, _, _ = stack(ref)
    ob_o, _, _ = stack(ours)
    _, _, dg_r = stack(ref2)
    _, _, dg_o = stack(ours2)
    assert same(ob_r.reset(), ob_o.reset()) and same(dg_r.reset(), dg_o.reset())
    rs = np.random.RandomState(1)
    logs = 0
    for t in range(50):
        act = rs.randn(n_agents, 2) * 2.0
        o1, r1, d1, i1 = ob_r.step(act)
        o2, r2, d2, i2 = ob_o.step(act)
        assert same(o1, o2) and same(r1, r2) and d1 == d2 and i1 == i2, t
        o1, r1, d1, l1 = dg_r.step(act)
        o2, r2, d2, l2 = dg_o.step(act)
        assert same(o1, o2) and same(r1, r2) and d1 == d2, t
        l1.pop('global/episode_time', None), l2.pop('global/episode_time', None)
        assert sorted(l1) == sorted(l2) and all(same(l1[k], l2[k]) for k in l1), (t, l1, l2)
        logs += bool(l1)
        if d1:
            assert same(dg_r.reset(), dg_o.reset()) and same(ob_r.reset(), ob_o.reset())
    assert logs >= 2


# ------------------------------------------------------------------------------------ rltools
class ScriptedPolicy(object):
    """What decrollout / dec_rollout need from a policy: per-agent actions from a seeded stream."""

    def __init__(self, n_agents, discrete, seed=3):
        self.n, self.discrete, self.seed = n_agents, discrete, seed

    def reset(self, dones=None):
        self.rs = np.random.RandomState(self.seed)

    def _draw(self):
        if self.discrete:
            return self.rs.randint(0, 5, size=(self.n, 1))
        return self.rs.randn(self.n, 2) * 2.0

    def sample_actions(self, obs):                  # rltools
        assert np.asarray(obs).shape[0] == self.n
        a = self._draw()
        return list(a), list(a.astype(np.float64))

    def get_actions(self, olist):                   # rllab
        assert len(olist) == self.n
        a = self._draw()
        return (list(a[:, 0]) if self.discrete else list(a)), dict(mean=a.astype(np.float64))


@pytest.mark.parametrize("family", ["ww", "hw", "pe"])
def test_rltools_decrollout_over_dropin_equals_over_reference(world, family):
    from oracle.refshim import load_rltools_samplers
    samplers = load_rltools_samplers()
    trajs = []
    for side in (0, 1):      # the Pursuit reference env shares a module-level stream proxy: build one at a time
        env = (make_pursuit_pair(world) if family == "pe" else make_pair(world, family))[side]
        pol = ScriptedPolicy(len(env.agents), discrete=(family == "pe"))
        trajs.append(samplers.decrollout(env, pol, 40, env.agents[0].action_space))
    a, b = trajs
    assert len(a) == len(b) >= 3
    # Pursuit contract: observations identical after casting the reference's float64 to
    # float32 -- the reference mixes float64 0.1 (out-of-bounds fill) with widened float32(0.1) cells
    cast = (lambda x: np.asarray(x, np.float64).astype(np.float32)) if family == "pe" else (lambda x: x)
    for ta, tb in zip(a, b):
        assert same(cast(ta.obs_T_Do), cast(tb.obs_T_Do), 0.0 if family == "pe" else 1e-9)
        assert same(ta.a_T_Da, tb.a_T_Da) and same(cast(ta.r_T), cast(tb.r_T))
        assert sorted(ta.info_D) == sorted(tb.info_D) and all(same(ta.info_D[k], tb.info_D[k]) for k in ta.info_D)
        assert len(ta) >= 10


# ------------------------------------------------------------------------------------ rllab
@pytest.mark.parametrize("family", ["ww", "hw"])
def test_rllab_dec_rollout_through_rllabenv(world, family):
    from oracle.refshim import load_rllab_callers
    RLLabEnv, ma_sampler, _ = load_rllab_callers()
    paths = []
    for env in make_pair(world, family):
        wrapped = RLLabEnv(env, ma_mode='decentralized')          # reads agents[0] spaces, timestep_limit
        assert wrapped.horizon == 1000 and wrapped.observation_space.flat_dim == env.agents[0].observation_space.shape[0]
        pol = ScriptedPolicy(len(env.agents), discrete=False)
        paths.append(ma_sampler.dec_rollout(wrapped, pol, max_path_length=30))
    for pa, pb in zip(*paths):
        for k in ('observations', 'actions', 'rewards'):
            assert same(pa[k], pb[k]), k
        assert sorted(pa['env_infos']) == sorted(pb['env_infos'])
        assert all(same(pa['env_infos'][k], pb['env_infos'][k]) for k in pa['env_infos'])
        assert pa['observations'].shape[0] == 30


@pytest.mark.parametrize("family", ["ww", "hw"])
def test_vec_env_executor_hook_equals_reference_vec_env_executor(world, family):
    """rllab's VecEnvExecutor over n reference envs (stream keys (SEED, 0..n-1)) against the batched
    executor the drop-in class hands out through `env.vec_env_executor(n, max_path_length)`: same
    obs lists, rewards, dones (horizon cut-off + reset in place) and stacked env_infos."""
    from oracle.refshim import load_rllab_callers
    RLLabEnv, _, VecEnvExecutor = load_rllab_callers()
    n, mpl = 3, 7
    ref_envs = [RLLabEnv(make_pair(world, family, env_id=i)[0], ma_mode='decentralized') for i in range(n)]
    ref_ex = VecEnvExecutor(ref_envs, mpl)
    ours = make_pair(world, family)[1]
    assert ours.vectorized is True
    our_ex = ours.vec_env_executor(n_envs=n, max_path_length=mpl)
    assert our_ex.num_envs == ref_ex.num_envs == n
    assert our_ex.observation_space.shape == ours.agents[0].observation_space.shape
    assert same(ref_ex.reset(), our_ex.reset())
    rs = np.random.RandomState(2)
    n_agents = len(ours.agents)
    resets = 0
    for t in range(20):
        act = rs.randn(n, n_agents, 2) * 2.0
        o1, r1, d1, i1 = ref_ex.step(list(act))
        o2, r2, d2, i2 = our_ex.step(act)
        assert same(o1, o2) and same(r1, r2) and np.array_equal(d1, d2), t
        assert sorted(i1) == sorted(i2) and all(np.array_equal(i1[k], i2[k]) for k in i1), t
        resets += int(np.sum(d1))
    assert resets >= 2 * n


# ------------------------------------------------------------------------------------ heuristic policies
def test_reference_policies_drive_the_dropin_envs(world):
    """The reference's hand-written policy OBJECTS (heuristics/waterworld.py, heuristics/pursuit.py) stepping the
    drop-in env classes, per agent as the reference's Visualizer does, reproduce the closed-loop goldens that the same
    policies produced on the reference envs (tests/golden/cl_*.npz)."""
    import json
    import os
    from conftest import GOLDEN_DIR, ROOT
    from oracle.heuristics_oracle import policy_draw
    from oracle.refshim import load_reference_heuristics
    WPol, PPol, _ = load_reference_heuristics()
    OurWW, OurPE, _ = world["ours"]
    g = np.load(os.path.join(GOLDEN_DIR, "cl_ww_c2.npz"))
    env = OurWW(seed=int(g["seed"]), env_id=int(g["env_id"]), dtype=torch.float64, **json.loads(str(g["config"])))
    pol = WPol(env.agents[0].observation_space, env.agents[0].action_space)
    o = env.reset()
    assert np.abs(np.array(o) - g["obs0"]).max() < 1e-9
    for t in range(60):
        a = np.array([pol.sample_actions(np.asarray(oi)[None])[0][0] for oi in o])
        assert np.abs(a - g["actions"][t]).max() < 1e-9, t
        o, r, d, i = env.step(a)
        assert np.abs(np.array(o) - g["obs"][t]).max() < 1e-9 and np.abs(r - g["rew"][t]).max() < 1e-9
        assert [i['evcatches'], i['pocatches']] == list(g["info"][t])
    # Pursuit (heuristics/pursuit.py as it runs under this interpreter: Python 3 division)
    g = np.load(os.path.join(GOLDEN_DIR, "cl_pe_conv_py3.npz"))
    maps = np.load(os.path.join(ROOT, "maps", "map_pool16.npy"))
    seed, env_id = int(g["seed"]), int(g["env_id"])
    env = OurPE(maps, seed=seed, env_id=env_id, **json.loads(str(g["config"])))

    class Space(object):
        def sample(self):
            raise AssertionError("every pursuer of this golden sees an evader")

    pol = PPol(None, Space())
    o = env.reset()
    assert np.array_equal(np.array(o, dtype=np.float32), np.asarray(g["obs0"], dtype=np.float32))
    for t in range(40):
        a = [int(pol.sample_actions(np.asarray(oq))[0]) for oq in o]
        assert a == list(g["actions"][t]), t
        o, r, d, i = env.step(a)
        assert np.array_equal(np.array(o, dtype=np.float32), np.asarray(g["obs"][t], dtype=np.float32))
        assert i['removed'] == g["removed"][t] and d == bool(g["done"][t])
    assert policy_draw(seed, env_id, 0, 0) in range(5)
