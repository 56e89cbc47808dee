"""Device heuristics vs a NumPy restatement of heuristics/waterworld.py:26-50 and
heuristics/pursuit.py:18-50 (runs on CPU tensors; the same code runs on CUDA tensors)."""
import math

import numpy as np
import torch

from madrl_b200.heuristics import pursuit_heuristic, waterworld_heuristic


def ww_ref(o, K):
    ang = np.linspace(0., 2. * np.pi, K + 1)[:-1]
    vecs = np.c_[np.cos(ang), np.sin(ang)]
    o = o[None]
    ob = -np.sum(o[:, 0:K][..., None] * vecs[None], axis=1)
    ev = np.sum(o[:, K:2 * K][..., None] * vecs[None], axis=1)
    po = -np.sum(o[:, 3 * K:4 * K][..., None] * vecs[None], axis=1)
    pu = np.sum(o[:, 5 * K:6 * K][..., None] * vecs[None], axis=1) / 2
    ev[o[:, 7 * K] > 0] *= 1.5
    po[o[:, 7 * K + 1] > 0] *= 1.5
    a = ob + ev + po + pu
    n = np.linalg.norm(a)
    return (a / n if n > 0 else np.zeros((1, 2)))[0]


def pe_ref(o):
    R = o.shape[0]
    x, y = R / 2, R / 2
    if np.sum(o[..., 2]) > 0:
        xev, yev = np.nonzero(o[..., 2])
        d = np.sqrt((xev - x) ** 2 + (yev - y) ** 2)
        k = np.argmin(d)
        xc, yc = xev[k], yev[k]
        if xc == x and yc == y:
            return 4
        ang = math.atan2(yc - y, xc - x)
        ang = (ang + np.pi) % (2 * np.pi) - np.pi
        if -np.pi / 4 <= ang < np.pi / 4:
            return 1
        if np.pi / 4 <= ang < 3 / 4. * np.pi:
            return 2
        if ang >= 3 / 4. * np.pi or ang < -3 / 4. * np.pi:
            return 0
        return 3
    return None


def test_waterworld_heuristic_matches_reference_formula():
    rs = np.random.RandomState(0)
    K = 30
    obs = rs.rand(64, 7 * K + 3) * (rs.rand(64, 7 * K + 3) < 0.2)
    obs[:, 7 * K:7 * K + 2] = rs.rand(64, 2) < 0.3
    obs[5] = 0
    got = waterworld_heuristic(torch.as_tensor(obs), K).numpy()
    for i in range(64):
        assert np.abs(got[i] - ww_ref(obs[i], K)).max() < 1e-12, i


def test_pursuit_heuristic_matches_reference_formula():
    rs = np.random.RandomState(1)
    for R in (7, 4):
        obs = np.zeros((200, R, R, 4))
        obs[..., 2] = (rs.rand(200, R, R) < 0.05) * 0.1
        got = pursuit_heuristic(torch.as_tensor(obs)).numpy()
        for i in range(200):
            want = pe_ref(obs[i])
            if want is None:
                assert 0 <= got[i] <= 4
            else:
                assert got[i] == want, (R, i)
