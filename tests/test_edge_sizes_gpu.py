"""Smallest and largest supported shapes of the three engines against the oracles (needs a GPU)."""
import numpy as np
import pytest
import torch

from oracle.hostage_oracle import HostageOracle
from oracle.philox import Stream
from oracle.pursuit_oracle import PursuitOracle
from oracle.waterworld_oracle import WaterworldOracle

pytestmark = pytest.mark.gpu


def f32(x):
    return np.asarray(x, dtype=np.float64).astype(np.float32)


@pytest.mark.parametrize("cfg,E,T", [
    (dict(n_pursuers=1, n_evaders=1, n_poison=1, n_sensors=1, n_coop=1, radius=0.05), 5, 60),   # minimum
    (dict(n_pursuers=32, n_evaders=3, n_poison=2, n_sensors=64, n_coop=3, radius=0.03), 2, 12),  # max pursuers / sensors
    (dict(n_pursuers=2, n_evaders=33, n_poison=30, n_sensors=33, n_coop=2, radius=0.03), 3, 25),  # 65 objects, 33 sensors
])
def test_waterworld_extreme_shapes_fp64(cfg, E, T):
    from madrl_b200 import BatchedMAWaterWorld
    eng = BatchedMAWaterWorld(E, dtype=torch.float64, seed=4, **cfg)
    obs0 = eng.reset().cpu().numpy()
    Np = cfg['n_pursuers']
    act = np.random.RandomState(1).randn(T, E, Np, 2) * 0.8
    obs, rew, done, info = [x.cpu().numpy() for x in eng.rollout(torch.as_tensor(act), auto_reset=False)]
    for e in range(E):
        o = WaterworldOracle(rng=Stream(4, e), **cfg)
        assert np.abs(np.array(o.reset()) - obs0[e]).max() < 1e-9
        for t in range(T):
            oo, rr, dd, ii = o.step(act[t, e])
            assert [ii['evcatches'], ii['pocatches']] == list(info[t, e]), (t, e)
            assert np.abs(np.array(oo) - obs[t, e]).max() < 1e-9 and np.abs(rr - rew[t, e]).max() < 1e-9


@pytest.mark.parametrize("maps,cfg,E,T", [
    (np.zeros((1, 2, 2), np.int32), dict(n_evaders=1, n_pursuers=1, obs_range=1, surround=False, n_catch=1,
                                         reward_mech='local', catchr=0.1), 6, 40),              # minimum
    (np.zeros((1, 2, 3), np.int32), dict(n_evaders=2, n_pursuers=2, obs_range=3, surround=True,
                                         reward_mech='global', catchr=0.1), 6, 60),
    (None, dict(n_evaders=64, n_pursuers=32, obs_range=11, surround=True, reward_mech='global', catchr=0.01,
                sample_maps=True), 2, 25),                                                        # maximum
])
def test_pursuit_extreme_shapes_bit_exact(maps, cfg, E, T):
    import os
    from conftest import ROOT
    from madrl_b200 import BatchedPursuitEvade
    if maps is None:
        maps = np.load(os.path.join(ROOT, "maps", "map_pool16.npy"))
    eng = BatchedPursuitEvade(E, maps, seed=6, **cfg)
    obs0 = eng.reset().cpu().numpy()
    Np = cfg['n_pursuers']
    act = np.random.RandomState(2).randint(0, 5, size=(T, E, Np)).astype(np.int32)
    obs, rew, done, removed = [x.cpu().numpy() for x in eng.rollout(torch.as_tensor(act), auto_reset=False)]
    for e in range(E):
        o = PursuitOracle(maps, rng=Stream(6, e), **cfg)
        assert np.array_equal(f32(o.reset()), obs0[e])
        for t in range(T):
            oo, rr, dd, ii = o.step(act[t, e])
            assert ii['removed'] == removed[t, e] and dd == bool(done[t, e]), (t, e)
            assert np.array_equal(f32(oo), obs[t, e]) and np.array_equal(f32(rr), rew[t, e]), (t, e)


@pytest.mark.parametrize("args,kw,E,T", [
    ((1, 1, 1, 1, 1), dict(n_sensors=1, radius=0.05, key_radius=0.05), 6, 80),                 # minimum
    ((32, 40, 30, 3, 2), dict(n_sensors=64, radius=0.03, key_radius=0.04, reward_mech='local'), 2, 20),
])
def test_hostage_extreme_shapes_fp64(args, kw, E, T):
    from madrl_b200 import BatchedHostageWorld
    eng = BatchedHostageWorld(E, *args, dtype=torch.float64, seed=8, **kw)
    obs0 = eng.reset().cpu().numpy()
    act = np.random.RandomState(3).randn(T, E, args[0], 2) * 2.0
    obs, rew, done, info = [x.cpu().numpy() for x in eng.rollout(torch.as_tensor(act), auto_reset=True)]
    for e in range(E):
        o = HostageOracle(*args, rng=Stream(8, e), **kw)
        assert np.abs(np.array(o.reset()) - obs0[e]).max() < 1e-9
        for t in range(T):
            oo, rr, dd, ii = o.step(act[t, e])
            assert [ii['ho_saved'], ii['cr_encs']] == list(info[t, e]) and dd == bool(done[t, e]), (t, e)
            if dd:
                oo = o.reset()
            assert np.abs(np.array(oo) - obs[t, e]).max() < 1e-9 and np.abs(rr - rew[t, e]).max() < 1e-9
