"""Parity of the CUDA ContinuousHostageWorld engine with the oracle / golden vectors (needs a
GPU): fp64 verification build on whole trajectories (identical discrete events, obs within 1e-9);
fp32 production build single-step teacher-forced within 1e-5 on non-fragile transitions."""
import json
import os

import numpy as np
import pytest
import torch

import teacher_forced as TF
from conftest import GOLDEN_DIR, ROOT
from oracle.hostage_oracle import HostageOracle
from oracle.philox import Stream

pytestmark = pytest.mark.gpu

TOL32 = TF.TOL32


def make(args, kw, E, dtype, **extra):
    from madrl_b200 import BatchedHostageWorld
    return BatchedHostageWorld(E, *args, dtype=dtype, **kw, **extra)


def engine_state(eng, e):
    st = {k: v.cpu().numpy() for k, v in eng.state.items()}
    Nr, Nc = eng.n_good, eng.n_bad
    X = np.stack([st['pos_x'][e], st['pos_y'][e]], 1).astype(np.float64)
    V = np.stack([st['vel_x'][e], st['vel_y'][e]], 1).astype(np.float64)
    f = int(st['flags'][e])
    return dict(rx=X[:Nr], rv=V[:Nr], cx=X[Nr:Nr + Nc], cv=V[Nr:Nr + Nc], hx=X[Nr + Nc:],
                key=st['key'][e].astype(np.float64)[None], bomb=st['bomb'][e].astype(np.float64)[None],
                saved=st['saved'][e].astype(bool), gate_open=bool(f & 1), bombed=bool(f & 2),
                t=int(st['timestep'][e]), counter=int(st['rng_counter'][e]))


CASES = {
    "c5": ((10, 16, 16, 4, 2), {}),
    "c5_local": ((10, 16, 16, 4, 2), dict(reward_mech='local')),
    "dense": ((3, 10, 5, 1, 2), dict(radius=0.05, sensor_range=0.35, key_radius=0.06, reward_mech='local', addid=False)),
    "k12_fixed_key": ((4, 6, 8, 2, 1), dict(radius=0.04, n_sensors=12, key_radius=0.05, bomb_radius=0.02,
                                            key_loc=np.array([[0.93, 0.97]]))),
    "big": ((12, 40, 30, 2, 2), dict(radius=0.03, n_sensors=40, key_radius=0.04)),
}


@pytest.mark.parametrize("name,E,T,std", [("c5", 24, 150, 1.0), ("c5_local", 16, 150, 2.0), ("dense", 32, 300, 3.0),
                                          ("k12_fixed_key", 32, 300, 3.0), ("big", 6, 60, 2.0)])
def test_fp64_trajectories_match_oracle(name, E, T, std):
    args, kw = CASES[name]
    seed, base = 321, 77
    eng = make(args, kw, E, torch.float64, seed=seed, env_id_base=base)
    obs0 = eng.reset().cpu().numpy()
    oracles = [HostageOracle(*args, rng=Stream(seed, base + e), **kw) for e in range(E)]
    for e, o in enumerate(oracles):
        assert np.abs(np.array(o.reset()) - obs0[e]).max() < 1e-9, e
    Nr = args[0]
    act = np.random.RandomState(5).randn(T, E, Nr, 2) * std
    obs, rew, done, info = [x.cpu().numpy() for x in eng.rollout(torch.as_tensor(act), auto_reset=True)]
    events = np.zeros(2, int)
    dones = 0
    for t in range(T):
        for e, o in enumerate(oracles):
            oo, rr, dd, ii = o.step(act[t, e])
            assert [ii['ho_saved'], ii['cr_encs']] == list(info[t, e]), (t, e)
            assert bool(done[t, e]) == dd, (t, e)
            assert np.abs(rr - rew[t, e]).max() < 1e-9, (t, e)
            if dd:                       # auto-reset: the slot holds the reset observation
                oo = o.reset()
                dones += 1
            assert np.abs(np.array(oo) - obs[t, e]).max() < 1e-9, (t, e)
            events += [ii['ho_saved'], ii['cr_encs']]
    for e, o in enumerate(oracles):
        s = engine_state(eng, e)
        assert s['counter'] == o.np_random.counter and s['t'] == o.t
        assert s['gate_open'] == o.gate_open and np.array_equal(s['saved'], o.saved)
        assert np.abs(s['cx'] - o.cx).max() < 1e-9 and np.abs(s['key'] - o.key_loc).max() < 1e-12
    assert events[1] > 0
    if name in ("dense", "k12_fixed_key"):
        assert events[0] > 0 and dones > 0


@pytest.mark.parametrize("name", ["hw_c5", "hw_c5_local", "hw_dense", "hw_k12"])
def test_fp64_matches_reference_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    kw = json.loads(str(g["config"]))
    if 'key_loc' in kw:
        kw['key_loc'] = np.array(kw['key_loc'])
    args = tuple(int(a) for a in g["args"])
    eng = make(args, kw, 1, torch.float64, seed=int(g["seed"]), env_id_base=int(g["env_id"]))
    assert np.abs(eng.reset().cpu().numpy()[0] - g["obs0"]).max() < 1e-9
    act = torch.as_tensor(g["actions"][:, None])
    obs, rew, done, info = [x.cpu().numpy() for x in eng.rollout(act, auto_reset=True)]
    assert np.array_equal(info[:, 0], g["info"])
    assert np.array_equal(done[:, 0].astype(bool), g["done"])
    assert np.abs(rew[:, 0] - g["rew"]).max() < 1e-9
    expect = g["obs"].copy()
    for k, t in enumerate(g["reset_at"]):          # the reference driver reset() where done
        expect[t] = g["reset_obs"][k]
    assert np.abs(obs[:, 0] - expect).max() < 1e-9
    assert int(eng.state['rng_counter'][0].item()) == int(g["counter"])


@pytest.mark.parametrize("name,E,T,std,min_frac", [("c5", 128, 16, 1.0, 0.98), ("dense", 128, 30, 3.0, 0.98)])
def test_fp32_single_step_teacher_forced(name, E, T, std, min_frac):
    """fp32 production build from its own states vs the float64 oracle, per-predicate exclusion
    (oracle/fragility.py); finished envs are re-initialised so that live envs keep being stepped."""
    args, kw = CASES[name]
    eng = make(args, kw, E, torch.float32, seed=99)
    eng.reset()
    log = TF.hw_self_teacher_forced(TF.TorchAdapter(eng), args, kw, 99, T, std, "hw_fp32_self_" + name,
                                    lambda done: eng.reset(mask=torch.as_tensor(done)))
    log.dump(ROOT)
    assert log.checked_frac >= min_frac and log.obs_frac >= 0.999, log.d


@pytest.mark.parametrize("name", ["hw_c5", "hw_c5_local", "hw_dense", "hw_k12"])
def test_fp32_teacher_forced_from_reference_states(name):
    """Every step of every golden replayed from the REAL reference's recorded float64 state (cast to
    fp32) and compared with the reference's recorded outputs."""
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    kw = json.loads(str(g["config"]))
    if 'key_loc' in kw:
        kw['key_loc'] = np.array(kw['key_loc'])
    args = tuple(int(a) for a in g["args"])
    eng = make(args, kw, 1, torch.float32, seed=int(g["seed"]), env_id_base=int(g["env_id"]))
    eng.reset()
    log = TF.hw_golden_teacher_forced(TF.TorchAdapter(eng), g, args, kw, "hw_fp32_golden_" + name)
    log.dump(ROOT)
    assert log.checked_frac >= 0.98 and log.obs_frac >= 0.999, log.d


def test_horizon_without_penalty_and_sharding():
    args, kw = CASES["c5"]
    E, T, mpl = 32, 12, 5
    eng = make(args, kw, E, torch.float32, seed=1, max_path_length=mpl)
    eng.reset()
    act = torch.zeros(T, E, 10, 2)
    obs, rew, done, info = [x.cpu() for x in eng.rollout(act, auto_reset=True)]
    # the executor horizon forces done without the env's not-saved penalty (hw:425-426 needs is_terminal)
    assert done[mpl - 1].all() and (rew[mpl - 1] > -40.0).all()   # penalty would be -3 * 16 = -48
    half = E // 2
    sh = make(args, kw, half, torch.float32, seed=1, env_id_base=half, max_path_length=mpl)
    sh.reset()
    o2 = sh.rollout(act[:, half:], auto_reset=True)
    assert torch.equal(o2[0].cpu(), obs[:, half:]) and torch.equal(o2[1].cpu(), rew[:, half:])


def test_dropin_env_surface():
    import pickle
    from madrl_b200 import ContinuousHostageWorld
    env = ContinuousHostageWorld(10, 16, 16, 4, 2, seed=3, env_id=1)
    assert len(env.agents) == 10 and env.agents[0].observation_space.shape == (156,)
    assert env.reward_mech == 'global' and env.timestep_limit == 1000
    obs = env.reset()
    orc = HostageOracle(10, 16, 16, 4, 2, rng=Stream(3, 1))
    assert np.abs(np.array(orc.reset()) - np.array(obs)).max() <= TOL32
    o1, r1, d1, i1 = env.step(np.zeros(20))
    assert len(o1) == 10 and r1.shape == (10,) and set(i1) == {'ho_saved', 'cr_encs'} and not env.is_gate_open
    env2 = pickle.loads(pickle.dumps(env))
    assert len(env2.reset()) == 10
    ex = env.vec_env_executor(n_envs=3, max_path_length=4)
    assert len(ex.reset()) == 3
    for _ in range(4):
        obs_n, rew_n, done_n, infos = ex.step(np.zeros((3, 20)))
    assert done_n.all() and rew_n.shape == (3, 10) and infos['cr_encs'].shape == (3,)
