"""Bit-exact parity of the CUDA PursuitEvade engine with the oracle and with the golden vectors
recorded from the real reference (needs a GPU).  Integer state, capture decisions, `removed`,
`done` identical; observations identical to the reference's float64 obs cast to float32;
rewards identical to the float64 reward narrowed to float32."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, ROOT
from oracle.philox import Stream
from oracle.pursuit_oracle import PursuitOracle

pytestmark = pytest.mark.gpu


def pool16():
    return np.load(os.path.join(ROOT, "maps", "map_pool16.npy"))


def small_map():
    m = np.zeros((1, 5, 5), dtype=np.int32)
    m[0, 2, 2] = -1
    return m


def make(maps, cfg, E, **kw):
    from madrl_b200 import BatchedPursuitEvade
    return BatchedPursuitEvade(E, maps, **cfg, **kw)


def f32(x):
    return np.asarray(x, dtype=np.float64).astype(np.float32)


C3 = dict(n_evaders=30, n_pursuers=8, obs_range=7, surround=True, n_catch=2, flatten=True,
          reward_mech='local', catchr=0.1, term_pursuit=5.0, sample_maps=True, include_id=True)
CASES = {
    "c3": (pool16, C3),
    "c3_global": (pool16, dict(C3, reward_mech='global', urgency_reward=-0.1)),
    "ncatch": (pool16, dict(C3, surround=False, n_evaders=20, n_pursuers=12, obs_range=5)),
    "window_r9": (pool16, dict(C3, constraint_window=0.5, n_evaders=6, n_pursuers=10, obs_range=9,
                               include_id=False)),
    "many_evaders": (pool16, dict(C3, n_evaders=50, n_pursuers=30, obs_range=11, catchr=0.01)),
    "crowd": (small_map, dict(n_evaders=4, n_pursuers=10, obs_range=3, surround=True, reward_mech='local',
                              catchr=0.1, term_pursuit=5.0)),
    "even_range": (small_map, dict(n_evaders=3, n_pursuers=3, obs_range=4, surround=False, n_catch=1,
                                   reward_mech='global')),
    "conv_layout": (pool16, dict(C3, flatten=False, n_evaders=12, obs_range=5)),     # (R, R, 4) obs
    "conv_small": (small_map, dict(n_evaders=3, n_pursuers=3, obs_range=4, surround=False, n_catch=1,
                                   reward_mech='global', flatten=False)),
    # random_opponents (pursuit_evade.py:81-82,177-181): 1..max_opponents-1 evaders per episode
    "random_opp": (small_map, dict(n_evaders=5, n_pursuers=6, obs_range=3, surround=False, n_catch=1,
                                   reward_mech='local', catchr=0.1, random_opponents=True, max_opponents=6)),
    "random_opp_c3": (pool16, dict(C3, random_opponents=True, max_opponents=31)),
}


def check_state(eng, oracles):
    st = {k: v.cpu().numpy() for k, v in eng.state.items()}
    for e, o in enumerate(oracles):
        assert np.array_equal(st['pursuer_x'][e], o.ppos[:, 0]) and np.array_equal(st['pursuer_y'][e], o.ppos[:, 1])
        live = ~o.gone
        assert np.array_equal(st['evader_x'][e][live], o.epos[live, 0])
        assert np.array_equal(st['evader_y'][e][live], o.epos[live, 1])
        gone_bits = [(int(st['gone'][e]) >> j) & 1 for j in range(o.Ne)]
        assert gone_bits == [int(x) for x in o.gone]
        assert int(st['rng_counter'][e]) == o.rng.counter
        # the never-cleared local_obs channels 1-2 (pe:119,438) as counts
        stale = st['stale'][e].astype(np.uint16)
        ln = np.float32(o.layer_norm)
        c1 = (stale & 0xff).astype(np.float32) / ln
        c2 = (stale >> 8).astype(np.float32) / ln
        assert np.array_equal(c1.reshape(o.Np, o.R, o.R), o.local_obs[:, 1].astype(np.float32))
        assert np.array_equal(c2.reshape(o.Np, o.R, o.R), o.local_obs[:, 2].astype(np.float32))


@pytest.mark.parametrize("name,E,T", [("c3", 24, 120), ("c3_global", 8, 80), ("ncatch", 16, 120),
                                      ("window_r9", 16, 120), ("many_evaders", 4, 40),
                                      ("crowd", 48, 300), ("even_range", 32, 150),
                                      ("conv_layout", 12, 100), ("conv_small", 16, 150),
                                      ("random_opp", 48, 200), ("random_opp_c3", 16, 60)])
def test_trajectories_bit_exact(name, E, T):
    mk, cfg = CASES[name]
    maps = mk()
    seed, base = 77, 500
    eng = make(maps, cfg, E, seed=seed, env_id_base=base)
    obs0 = eng.reset().cpu().numpy()
    oracles = [PursuitOracle(maps, rng=Stream(seed, base + e), **cfg) for e in range(E)]
    for e, o in enumerate(oracles):
        assert np.array_equal(f32(o.reset()).reshape(obs0[e].shape), obs0[e]), e
    check_state(eng, oracles)
    Np = cfg['n_pursuers']
    act = np.random.RandomState(9).randint(0, 5, size=(T, E, Np)).astype(np.int32)
    obs, rew, done, removed = [x.cpu().numpy() for x in eng.rollout(torch.as_tensor(act), auto_reset=False)]
    total_removed = 0
    for t in range(T):
        for e, o in enumerate(oracles):
            oo, rr, dd, ii = o.step(act[t, e])
            assert ii['removed'] == removed[t, e] and dd == bool(done[t, e]), (t, e)
            assert np.array_equal(f32(oo).reshape(obs[t, e].shape), obs[t, e]), (t, e)
            assert np.array_equal(f32(rr), rew[t, e]), (t, e, rr, rew[t, e])
            total_removed += ii['removed']
    check_state(eng, oracles)
    if name in ("ncatch", "crowd", "even_range"):
        assert total_removed > 0


@pytest.mark.parametrize("name", ["pe_c3", "pe_c3_global", "pe_ncatch", "pe_window", "pe_small",
                                  "pe_even_range", "pe_crowd", "pe_random_opp"])
def test_matches_reference_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    cfg = json.loads(str(g["config"]))
    maps = pool16() if str(g["maps"]) == "pool16" else small_map()
    eng = make(maps, cfg, 1, seed=int(g["seed"]), env_id_base=int(g["env_id"]))
    assert np.array_equal(eng.reset().cpu().numpy()[0], f32(g["obs0"]))
    resets = list(g["reset_at"])
    t0, k = 0, 0
    T = g["actions"].shape[0]
    while t0 < T:                      # roll segment by segment between the recorded reset() calls
        t1 = (resets[k] + 1) if k < len(resets) else T
        act = torch.as_tensor(g["actions"][t0:t1, None])
        obs, rew, done, removed = [x.cpu().numpy() for x in eng.rollout(act, auto_reset=False)]
        assert np.array_equal(obs[:, 0], f32(g["obs"][t0:t1]))
        assert np.array_equal(rew[:, 0], f32(g["rew"][t0:t1]))
        assert np.array_equal(done[:, 0].astype(bool), g["done"][t0:t1])
        assert np.array_equal(removed[:, 0], g["removed"][t0:t1])
        if k < len(resets):
            assert np.array_equal(eng.reset().cpu().numpy()[0], f32(g["reset_obs"][k]))
            k += 1
        t0 = t1
    assert int(eng.state['rng_counter'][0].item()) == int(g["counter"])


@pytest.mark.parametrize("name", ["crowd", "random_opp"])
def test_auto_reset_and_horizon(name):
    mk, cfg = CASES[name]
    maps = mk()
    E, T, mpl, seed = 12, 60, 17, 3
    eng = make(maps, cfg, E, seed=seed, max_path_length=mpl)
    eng.reset()
    Np = cfg['n_pursuers']
    act = np.random.RandomState(2).randint(0, 5, size=(T, E, Np)).astype(np.int32)
    obs, rew, done, removed = [x.cpu().numpy() for x in eng.rollout(torch.as_tensor(act), auto_reset=True)]
    oracles = [PursuitOracle(maps, rng=Stream(seed, e), **cfg) for e in range(E)]
    for o in oracles:
        o.reset()
    ts = np.zeros(E, int)
    n_done = 0
    for t in range(T):
        for e, o in enumerate(oracles):
            oo, rr, dd, ii = o.step(act[t, e])
            ts[e] += 1
            dd = dd or ts[e] >= mpl
            assert bool(done[t, e]) == dd and removed[t, e] == ii['removed']
            if dd:
                oo = o.reset()
                ts[e] = 0
                n_done += 1
            assert np.array_equal(f32(oo), obs[t, e]), (t, e)
            assert np.array_equal(f32(rr), rew[t, e])
    assert n_done >= E * (T // mpl)


def test_sharding_and_full_size_properties():
    """Full BASELINE size (65536 envs): size-independent properties instead of the slow oracle."""
    maps = pool16()
    E, T = 65536, 4
    eng = make(maps, C3, E, seed=11)
    obs0 = eng.reset()
    act = torch.randint(0, 5, (T, E, 8), dtype=torch.int32, device='cuda')
    obs, rew, done, removed = eng.rollout(act, auto_reset=True)
    st = eng.state
    m = torch.as_tensor(maps, device='cuda')
    # nobody stands in a building, everybody is on the map
    for xk, yk in (("pursuer_x", "pursuer_y"), ("evader_x", "evader_y")):
        x, y = st[xk].long(), st[yk].long()
        assert (x < 16).all() and (y < 16).all()
        assert (m[st['map_id'].long()[:, None], x, y] == 0).all()
    # observations only take values k/10 and the id channel is i/Np
    vals = torch.unique(obs)
    assert all(abs(float(v) * 10 - round(float(v) * 10)) < 1e-5 or abs(float(v) * 8 - round(float(v) * 8)) < 1e-6 for v in vals)
    assert torch.equal(obs[..., -1], (torch.arange(8, device='cuda') / 8.0).expand(T, E, 8).float())
    # window centre of channel 1 always shows the observing pursuer itself
    centre = obs[..., 49 + 3 * 7 + 3]
    assert (centre >= np.float32(0.1)).all()
    # env e of a shard with env_id_base reproduces env (base+e) of the full batch
    sh = make(maps, C3, 256, seed=11, env_id_base=4096)
    assert torch.equal(sh.reset(), obs0[4096:4096 + 256])
    o2 = sh.rollout(act[:, 4096:4096 + 256].contiguous(), auto_reset=True)
    assert torch.equal(o2[0], obs[:, 4096:4096 + 256]) and torch.equal(o2[1], rew[:, 4096:4096 + 256])


def test_dropin_env_surface():
    import pickle
    from madrl_b200 import PursuitEvade
    maps = pool16()
    env = PursuitEvade(maps, seed=5, env_id=1, **C3)
    assert len(env.agents) == 8 and env.agents[0].observation_space.shape == (148,)
    assert env.agents[0].action_space.n == 5 and env.reward_mech == 'local'
    obs = env.reset()
    orc = PursuitOracle(maps, rng=Stream(5, 1), **C3)
    assert np.array_equal(f32(orc.reset()), np.array(obs, dtype=np.float32))
    a = [0, 1, 2, 3, 4, 0, 1, 2]
    o1, r1, d1, i1 = env.step(a)
    o2, r2, d2, i2 = orc.step(a)
    assert np.array_equal(f32(o2), np.array(o1, dtype=np.float32)) and np.array_equal(f32(r2), r1.astype(np.float32))
    assert d1 == d2 and i1 == i2 and not env.is_terminal
    joint = int(np.ravel_multi_index(a, [5] * 8))        # joint scalar action (pe:233)
    o1, r1, d1, i1 = env.step(joint)
    o2, r2, d2, i2 = orc.step(joint)
    assert np.array_equal(f32(o2), np.array(o1, dtype=np.float32))
    env.catchr = 0.3                                     # curriculum fields survive pickling (pe:397-411)
    env2 = pickle.loads(pickle.dumps(env))
    assert env2.catchr == 0.3 and env2.n_pursuers == 8 and len(env2.reset()) == 8
    ex = env.vec_env_executor(n_envs=5, max_path_length=4)
    assert len(ex.reset()) == 5
    for _ in range(4):
        obs_n, rew_n, done_n, infos = ex.step(np.zeros((5, 8), dtype=int))
    assert done_n.all() and rew_n.shape == (5, 8) and infos['removed'].shape == (5,)
