"""Device-resident versions of the reference's hand-written policies: action
generators for benchmarks and smoke rollouts that never leave the GPU.  They are a handful of
dense tensor expressions over the observation batch, written with torch ops (plumbing, not a hot
path).

* ``waterworld_heuristic``  heuristics/waterworld.py:11-53 -- flee obstacles and poison, chase
  evaders, close in on allies; obs layout documented at heuristics/waterworld.py:12-22.
* ``pursuit_heuristic``     heuristics/pursuit.py:18-50 -- walk towards the nearest visible evader
  of the (R, R, 4) observation, random move otherwise.
"""
import math

import torch


def waterworld_heuristic(obs, n_sensors):
    """obs [..., 7K+2(+1)] -> actions [..., 2].

    The reference normalises by the Frobenius norm of the WHOLE batch it is called with
    (heuristics/waterworld.py:46-50); it is called per agent (B = 1) in the reference's own demo,
    so here every agent row is normalised by its own norm."""
    K = n_sensors
    ang = torch.arange(K, device=obs.device, dtype=obs.dtype) * (2.0 * math.pi / K)
    vecs = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1)             # [K, 2]
    ob = -(obs[..., 0:K].unsqueeze(-1) * vecs).sum(-2)
    ev = (obs[..., K:2 * K].unsqueeze(-1) * vecs).sum(-2)
    po = -(obs[..., 3 * K:4 * K].unsqueeze(-1) * vecs).sum(-2)
    pu = (obs[..., 5 * K:6 * K].unsqueeze(-1) * vecs).sum(-2) / 2
    ev = torch.where((obs[..., 7 * K] > 0).unsqueeze(-1), ev * 1.5, ev)
    po = torch.where((obs[..., 7 * K + 1] > 0).unsqueeze(-1), po * 1.5, po)
    act = ob + ev + po + pu
    norm = act.norm(dim=-1, keepdim=True)
    return torch.where(norm > 0, act / norm.clamp_min(1e-30), torch.zeros_like(act))


def pursuit_heuristic(obs, generator=None):
    """obs [..., R, R, 4] (flatten=False layout) -> int32 actions [...]
    (0 left, 1 right, 2 up, 3 down, 4 stay)."""
    R = obs.shape[-2]
    lead = obs.shape[:-3]
    ev = obs[..., 2].reshape(-1, R * R)
    B = ev.shape[0]
    cx = cy = R / 2                                                             # heuristics/pursuit.py:23
    idx = torch.arange(R * R, device=obs.device)
    xs, ys = (idx // R).to(obs.dtype), (idx % R).to(obs.dtype)
    d = torch.sqrt((xs - cx) ** 2 + (ys - cy) ** 2).expand(B, -1)
    d = torch.where(ev > 0, d, torch.full_like(d, float("inf")))
    # np.nonzero order = row-major; argmin keeps the first minimum
    j = torch.argmin(d, dim=1)
    seen = (ev > 0).any(dim=1)
    dx, dy = xs[j] - cx, ys[j] - cy
    ang = torch.atan2(dy, dx)
    ang = torch.remainder(ang + math.pi, 2 * math.pi) - math.pi
    a = torch.full((B,), 3, dtype=torch.int32, device=obs.device)              # down
    a = torch.where((ang >= -math.pi / 4) & (ang < math.pi / 4), torch.ones_like(a), a)          # right
    a = torch.where((ang >= math.pi / 4) & (ang < 0.75 * math.pi), torch.full_like(a, 2), a)    # up
    a = torch.where((ang >= 0.75 * math.pi) | (ang < -0.75 * math.pi), torch.zeros_like(a), a)  # left
    a = torch.where((dx == 0) & (dy == 0), torch.full_like(a, 4), a)                            # stay
    rnd = torch.randint(0, 5, (B,), dtype=torch.int32, device=obs.device, generator=generator)
    return torch.where(seen, a, rnd).reshape(lead)
