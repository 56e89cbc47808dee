"""The in-kernel heuristic policies on the GPU (needs a GPU): the same closed-loop checks as
tests/test_heuristics.py runs through the emulator -- the engine's actions are the oracle policy's on the
previous observation, and the env those actions drive is the oracle env (Waterworld fp64 <= 1e-9 over whole
trajectories, fp32 <= 1e-5 per decision; Pursuit bit-exact) -- plus full-size properties at the
BASELINE.json batch sizes, and the stand-alone generators of csrc/heuristics.cu."""
import numpy as np
import pytest
import torch

from oracle.heuristics_oracle import pursuit_action, waterworld_action
import json
import os

from conftest import GOLDEN_DIR
from test_heuristics import (CL_PE, CL_WW, PE, WW, check_closed_loop_golden_pe, check_closed_loop_golden_ww,
                             check_pursuit_closed_loop, check_waterworld_closed_loop, pool16, small_map)

pytestmark = pytest.mark.gpu


def n(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("name,E,T", [("c2", 16, 120), ("dense", 8, 100), ("c4", 3, 25), ("k40_randobst", 4, 40)])
def test_waterworld_policy_closes_the_loop_fp64(name, E, T):
    from madrl_b200 import BatchedMAWaterWorld
    cfg, seed, base, mpl = WW[name], 11, 300, 17
    eng = BatchedMAWaterWorld(E, seed=seed, env_id_base=base, max_path_length=mpl, dtype=torch.float64, **cfg)
    obs0 = eng.reset()
    out = eng.rollout_heuristic(T, obs0, auto_reset=True)
    act, obs, rew, done, info = [n(x) for x in out]
    checked, _ = check_waterworld_closed_loop(eng, cfg, seed, base, n(obs0), act, obs, rew, done, info, mpl, 1e-9, 0.0)
    assert checked == T * E * cfg['n_pursuers'] and done.any()
    # chunk invariance: two launches chained through the last observation == one launch
    eng2 = BatchedMAWaterWorld(E, seed=seed, env_id_base=base, max_path_length=mpl, dtype=torch.float64, **cfg)
    a1 = eng2.rollout_heuristic(T // 2, eng2.reset(), auto_reset=True)
    a2 = eng2.rollout_heuristic(T - T // 2, a1[1][-1].contiguous(), auto_reset=True)
    for x, y1, y2 in zip(out, a1, a2):
        assert torch.equal(x, torch.cat([y1, y2]))


@pytest.mark.parametrize("name,E,T", [("c2", 64, 60), ("c4", 6, 12)])
def test_waterworld_policy_fp32_decisions(name, E, T):
    """The production (fp32) instantiation: every decision within 1e-5 of the oracle policy evaluated on the
    engine's own previous observation (rows whose un-normalised sum is below 1e-2 amplify float32 rounding
    through the division and are skipped; counted)."""
    from madrl_b200 import BatchedMAWaterWorld
    cfg = WW[name]
    eng = BatchedMAWaterWorld(E, seed=5, env_id_base=40, **cfg)
    obs0 = eng.reset()
    act, obs, rew, done, info = [n(x) for x in eng.rollout_heuristic(T, obs0, auto_reset=False)]
    prev_all = np.concatenate([n(obs0)[None], obs[:-1]]).astype(np.float64)
    checked = skipped = 0
    for t in range(T):
        for e in range(E):
            for i in range(cfg['n_pursuers']):
                want, nrm = waterworld_action(prev_all[t, e, i], return_norm=True)
                if nrm == 0:
                    assert not act[t, e, i].any()
                elif nrm > 1e-2:
                    checked += 1
                    assert np.abs(want - act[t, e, i]).max() < 1e-5, (t, e, i)
                else:
                    skipped += 1
    assert checked > 0.5 * T * E * cfg['n_pursuers'] and skipped < 0.1 * checked


def test_waterworld_policy_full_batch_properties():
    """BASELINE configs[1] batch (4096 envs): unit or zero actions, zero exactly where nothing relevant is sensed,
    and the policy does its job: it catches more food and less poison than zero actions from the same start."""
    from madrl_b200 import BatchedMAWaterWorld
    E, T, cfg, K = 4096, 200, WW["c2"], 30
    eng = BatchedMAWaterWorld(E, seed=1, **cfg)
    obs0 = eng.reset()
    act, obs, rew, done, info = eng.rollout_heuristic(T, obs0, auto_reset=True)
    nrm = act.norm(dim=-1)
    assert bool(((nrm - 1).abs() < 1e-5).logical_or(nrm == 0).all())
    prev = torch.cat([obs0[None], obs[:-1]])
    sensed = (prev[..., 0:K].abs().sum(-1) + prev[..., K:2 * K].abs().sum(-1) + prev[..., 3 * K:4 * K].abs().sum(-1) +
              prev[..., 5 * K:6 * K].abs().sum(-1)) > 0
    assert bool((nrm[~sensed] == 0).all()) and float((nrm[sensed] > 0).float().mean()) > 0.999
    eng0 = BatchedMAWaterWorld(E, seed=1, **cfg)
    eng0.reset()
    _, _, _, info0 = eng0.rollout(torch.zeros(T, E, 5, 2, device=act.device), auto_reset=True)
    ev, po = info[..., 0].sum().item(), info[..., 1].sum().item()
    ev0, po0 = info0[..., 0].sum().item(), info0[..., 1].sum().item()
    assert ev > 2 * ev0 and po < po0, (ev, ev0, po, po0)


@pytest.mark.parametrize("py2", [True, False])
@pytest.mark.parametrize("name,E,T", [("c3", 8, 60), ("sparse_conv", 16, 80), ("even_r4", 16, 80), ("r9_global", 6, 50)])
def test_pursuit_policy_closes_the_loop(name, E, T, py2):
    from madrl_b200 import BatchedPursuitEvade
    mk, cfg = PE[name]
    maps, seed, base, mpl = mk(), 21, 900, 13
    eng = BatchedPursuitEvade(E, maps, seed=seed, env_id_base=base, max_path_length=mpl, **cfg)
    obs0 = eng.reset()
    out = eng.rollout_heuristic(T, obs0, auto_reset=True, py2_division=py2)
    act, obs, rew, done, removed = [n(x) for x in out]
    n_random = check_pursuit_closed_loop(maps, cfg, seed, base, n(obs0), act, obs, rew, done, removed, mpl, py2)
    assert done.any() and (name == "c3" or n_random > 0)
    eng2 = BatchedPursuitEvade(E, maps, seed=seed, env_id_base=base, max_path_length=mpl, **cfg)
    a1 = eng2.rollout_heuristic(T // 3, eng2.reset(), auto_reset=True, py2_division=py2)
    a2 = eng2.rollout_heuristic(T - T // 3, a1[1][-1].contiguous(), auto_reset=True, py2_division=py2)
    for x, y1, y2 in zip(out, a1, a2):
        assert torch.equal(x, torch.cat([y1, y2]))


def test_pursuit_policy_full_batch_properties():
    """BASELINE configs[2] batch (65 536 envs): valid actions, and the policy out-earns random moves."""
    from madrl_b200 import BatchedPursuitEvade
    mk, cfg = PE["c3"]
    E, T = 65536, 40
    eng = BatchedPursuitEvade(E, mk(), seed=2, **cfg)
    act, obs, rew, done, removed = eng.rollout_heuristic(T, eng.reset(), auto_reset=True)
    assert int(act.min()) >= 0 and int(act.max()) <= 4
    engr = BatchedPursuitEvade(E, mk(), seed=2, **cfg)
    engr.reset()
    _, rew_r, _, _ = engr.rollout(torch.randint(0, 5, (T, E, 8), dtype=torch.int32, device=act.device), auto_reset=True)
    # the policy walks onto the nearest evader: it collects the neighbouring-evader reward (pursuit_evade.py:359-381)
    # much faster than random moves (it does not try to surround, so it does not remove more evaders)
    assert rew.double().sum().item() > 1.3 * rew_r.double().sum().item()


def test_generators_match_oracle():
    from madrl_b200.heuristics import pursuit_heuristic, waterworld_heuristic
    rs = np.random.RandomState(3)
    K, D = 30, 213
    obs = rs.rand(500, D) * (rs.rand(500, D) < 0.2)
    obs[:, 7 * K:7 * K + 2] = rs.rand(500, 2) < 0.3
    obs[4] = 0
    for dt, tol in ((torch.float64, 1e-12), (torch.float32, 1e-5)):
        got = n(waterworld_heuristic(torch.as_tensor(obs, dtype=dt, device="cuda").view(50, 10, D), K)).reshape(500, 2)
        for i in range(500):
            want, nrm = waterworld_action(obs[i], return_norm=True)
            if nrm == 0 or nrm > 1e-2 or dt == torch.float64:
                assert np.abs(got[i] - want).max() < tol, i
    for R in (7, 4):
        m = 300
        conv = np.zeros((m, R, R, 4), np.float32)
        conv[..., 2] = (rs.rand(m, R, R) < 0.04) * 0.1
        fb = rs.randint(0, 5, m)
        for py2 in (True, False):
            want = [pursuit_action(conv[i], lambda i=i: int(fb[i]), py2) for i in range(m)]
            got = pursuit_heuristic(torch.as_tensor(conv, device="cuda"), py2_division=py2, fallback=torch.as_tensor(fb))
            assert list(n(got)) == want
            flat = np.concatenate([np.zeros((m, 2 * R * R), np.float32), conv[..., 2].reshape(m, -1), np.ones((m, 1), np.float32)], 1)
            got = pursuit_heuristic(torch.as_tensor(flat, device="cuda"), obs_range=R, py2_division=py2, fallback=torch.as_tensor(fb))
            assert list(n(got)) == want


@pytest.mark.parametrize("name", CL_WW)
def test_waterworld_policy_reproduces_reference_closed_loop(name):
    """The REAL reference env stepped by the REAL reference policy (tests/golden/cl_ww_*.npz) vs the fp64 engine's
    in-kernel policy rollout: actions, observations, rewards <= 1e-9, catches equal, over the whole trajectory."""
    from madrl_b200 import BatchedMAWaterWorld
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    eng = BatchedMAWaterWorld(1, seed=int(g["seed"]), env_id_base=int(g["env_id"]), dtype=torch.float64,
                              **json.loads(str(g["config"])))
    last = {}

    def reset():
        last["o"] = eng.reset()
        return n(last["o"])

    check_closed_loop_golden_ww(g, reset, lambda T, o: [n(x) for x in eng.rollout_heuristic(T, last["o"], auto_reset=False)], 1e-9)
    assert int(eng.state["rng_counter"][0]) == int(g["counter"])


@pytest.mark.parametrize("name", CL_PE)
def test_pursuit_policy_reproduces_reference_closed_loop(name):
    from madrl_b200 import BatchedPursuitEvade
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    maps = pool16() if str(g["maps"]) == "pool16" else small_map()
    eng = BatchedPursuitEvade(1, maps, seed=int(g["seed"]), env_id_base=int(g["env_id"]), **json.loads(str(g["config"])))
    last = {}

    def reset():
        last["o"] = eng.reset()
        return n(last["o"])

    check_closed_loop_golden_pe(g, reset, lambda T, o: [n(x) for x in eng.rollout_heuristic(
        T, last["o"], auto_reset=False, py2_division=bool(g["py2"]))])
    assert int(eng.state["rng_counter"][0]) == int(g["counter"])
