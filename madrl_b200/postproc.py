"""Device-side consumers of the rollout tensors: what the reference does
right after ``env.step`` in its sampler loop, as streaming passes over ``[T, E, A, ...]`` tensors.

* ``gae``            rllab ``BaseSampler.process_samples`` (rllab/rllab/sampler/base.py:48-68)
* ``FrameStack``     ``ObservationBuffer`` (madrl_environments/__init__.py:143-196)
* ``Standardizer``   ``StandardizedEnv``   (madrl_environments/__init__.py:204-291)
* ``center_advantages`` / ``explained_variance``  the whole-batch statistics of
                     ``process_samples`` (base.py:69-86; rllab/rllab/algos/util.py:7-12,
                     rllab/rllab/misc/special.py:51-59)
* ``to_paths``       the per-agent ``paths`` dicts of ``dec_rollout``
                     (rllab/rllab/sampler/ma_sampler.py:52-100), built on the host from gathered
                     tensors for callers that still want rllab's list-of-paths format.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def gae(rew, values, done, discount, gae_lambda, last_value=None):
    """rew, values [T,E,A] float32 cuda; done [T,E] uint8 -> (advantages, returns) [T,E,A].

    Paths end where ``done`` is set (bootstrap 0, base.py:57) and at the end of the rollout
    (bootstrap ``last_value`` [E,A] if given, else 0)."""
    rew, values, done = rew.contiguous(), values.contiguous(), done.contiguous()
    T, E, A = rew.shape
    adv, ret = torch.empty_like(rew), torch.empty_like(rew)
    lv = last_value.contiguous() if last_value is not None else None
    with torch.cuda.device(rew.device):
        _lib.check(_lib.lib().madrl_gae_f32(T, E, A, _ptr(rew), _ptr(values), _ptr(done), _ptr(lv),
                                            float(discount), float(gae_lambda), _ptr(adv), _ptr(ret),
                                            _stream(rew.device)))
    return adv, ret


MOMENTS_STATS, MOMENTS_WS = 9, 4096            # include/madrl_b200.h


def _moments_buffers(dev):
    return (torch.empty(MOMENTS_STATS, dtype=torch.float64, device=dev),
            torch.empty(MOMENTS_WS, dtype=torch.float64, device=dev))


def center_advantages(adv, center=True, positive=False, inplace=False):
    """``util.center_advantages`` / ``util.shift_advantages_to_positive`` over every sample of the
    batch (base.py:82-86 applies them to the concatenation of all paths): float64 mean / population
    std / min computed on the device, deterministic.  Returns the processed tensor (same shape)."""
    out = adv.contiguous() if inplace else adv.contiguous().clone()
    stats, ws = _moments_buffers(out.device)
    with torch.cuda.device(out.device):
        _lib.check(_lib.lib().madrl_center_advantages_f32(out.numel(), _ptr(out), int(bool(center)),
                                                          int(bool(positive)), _ptr(stats), _ptr(ws),
                                                          _stream(out.device)))
    return out


def moments(a, b=None):
    """float64 [9] cuda tensor: means, population variances and minima of the series a, b, b - a
    taken over every element (NumPy's two-pass ``var``)."""
    a = a.contiguous()
    b = b.contiguous() if b is not None else None
    assert b is None or b.numel() == a.numel()
    stats, ws = _moments_buffers(a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().madrl_moments_f32(a.numel(), _ptr(a), _ptr(b), _ptr(stats), _ptr(ws),
                                                _stream(a.device)))
    return stats


def explained_variance(ypred, y):
    """``special.explained_variance_1d`` (special.py:51-59) of baseline predictions against returns,
    over every sample of the batch; one 72-byte device->host read."""
    st = moments(ypred, y).cpu().numpy()
    var_pred, vary, var_res = float(st[3]), float(st[4]), float(st[5])
    if np.isclose(vary, 0):
        return 0 if var_pred > 0 else 1
    return 1 - var_res / (vary + 1e-8)


class FrameStack(object):
    """ObservationBuffer for a whole env batch: the last ``buffer_size`` observations of every
    agent stacked on a new minor axis; a reset fills every slot with the reset observation."""

    def __init__(self, n_envs, n_agents, obs_dim, buffer_size, device):
        self.shape = (n_envs, n_agents, obs_dim, buffer_size)
        self.carry = torch.zeros(self.shape, dtype=torch.float32, device=device)   # __init__.py:150

    def reset(self, obs0):
        """obs0 [E,A,D] from engine.reset() -> stacked [E,A,D,B] (__init__.py:186-196)."""
        self.carry.copy_(obs0.unsqueeze(-1).expand(self.shape))
        return self.carry.clone()

    def rollout(self, obs, done):
        """obs [T,E,A,D], done [T,E] from engine.rollout(auto_reset=True) -> [T,E,A,D,B]."""
        obs, done = obs.contiguous(), done.contiguous()
        T = obs.shape[0]
        E, A, D, B = self.shape
        out = torch.empty((T, E, A, D, B), dtype=torch.float32, device=obs.device)
        with torch.cuda.device(obs.device):
            _lib.check(_lib.lib().madrl_frame_stack_f32(T, E, A, D, B, _ptr(obs), _ptr(done),
                                                        _ptr(self.carry), _ptr(out), _stream(obs.device)))
        return out


class Standardizer(object):
    """StandardizedEnv's running observation / reward normalisation for a whole env batch
    (one running estimate per env, agent and observation coordinate, as each reference env keeps
    its own).  Deviation from the wrapped-single-env reference under auto-reset: the terminal
    observation of an episode is replaced by the reset observation in the rollout tensor, so the
    running estimate sees the reset observation only (the reference updates on both)."""

    def __init__(self, n_envs, n_agents, obs_dim, device, scale_reward=1., enable_obsnorm=False,
                 enable_rewnorm=False, obs_alpha=0.001, rew_alpha=0.001, eps=1e-8):
        self.scale_reward, self.enable_obsnorm, self.enable_rewnorm = scale_reward, enable_obsnorm, enable_rewnorm
        self.obs_alpha, self.rew_alpha, self.eps = obs_alpha, rew_alpha, eps
        self.obs_mean = torch.zeros((n_envs, n_agents, obs_dim), dtype=torch.float64, device=device)
        self.obs_var = torch.ones((n_envs, n_agents, obs_dim), dtype=torch.float64, device=device)
        self.rew_mean = torch.zeros((n_envs, n_agents), dtype=torch.float64, device=device)
        self.rew_var = torch.ones((n_envs, n_agents), dtype=torch.float64, device=device)

    def obs(self, obs):
        """In place on obs [T,E,A,D] (or [E,A,D])."""
        if not self.enable_obsnorm:
            return obs
        x = obs if obs.dim() == 4 else obs.unsqueeze(0)
        assert x.is_contiguous()
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().madrl_standardize_f32(x.shape[0], self.obs_mean.numel(), _ptr(x),
                                                        _ptr(self.obs_mean), _ptr(self.obs_var), self.obs_alpha,
                                                        self.eps, 1, 1.0, 1, _stream(x.device)))
        return obs

    def rew(self, rew):
        """In place on rew [T,E,A]: optional running-std division, then scale_reward."""
        assert rew.is_contiguous()
        with torch.cuda.device(rew.device):
            _lib.check(_lib.lib().madrl_standardize_f32(rew.shape[0], self.rew_mean.numel(), _ptr(rew),
                                                        _ptr(self.rew_mean), _ptr(self.rew_var), self.rew_alpha,
                                                        self.eps, 0, float(self.scale_reward),
                                                        int(self.enable_rewnorm), _stream(rew.device)))
        return rew


class EpisodeStats(object):
    """DiagnosticsWrapper's per-episode statistics for a whole env batch
    (madrl_environments/__init__.py:314-369): episode reward per agent, its agent-mean, the
    discounted return of the agent-mean reward and the episode length, emitted at the steps where
    an episode closes (``done`` or ``max_traj_len`` reached)."""

    def __init__(self, n_envs, n_agents, device, discount=0.99, max_traj_len=500):
        self.n_envs, self.n_agents, self.discount, self.max_traj_len = n_envs, n_agents, discount, max_traj_len
        self.carry = torch.zeros((n_envs, n_agents + 3), dtype=torch.float64, device=device)

    def rollout(self, rew, done):
        """rew [T,E,A] float32, done [T,E] uint8 -> dict(end [T,E] bool, episode_reward [T,E,A],
        episode_avg_reward [T,E], episode_disc_return [T,E], episode_length [T,E])."""
        rew, done = rew.contiguous(), done.contiguous()
        T, E, A = rew.shape
        dev = rew.device
        ep_r = torch.empty((T, E, A), dtype=torch.float32, device=dev)
        ep_d = torch.empty((T, E), dtype=torch.float32, device=dev)
        ep_l = torch.empty((T, E), dtype=torch.int32, device=dev)
        ep_e = torch.empty((T, E), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().madrl_episode_stats_f32(T, E, A, _ptr(rew), _ptr(done), float(self.discount),
                                                          int(self.max_traj_len), _ptr(self.carry), _ptr(ep_r),
                                                          _ptr(ep_d), _ptr(ep_l), _ptr(ep_e), _stream(dev)))
        return dict(end=ep_e.bool(), episode_reward=ep_r, episode_avg_reward=ep_r.mean(dim=-1),
                    episode_disc_return=ep_d, episode_length=ep_l)


def to_paths(obs, actions, rew, done, infos=None):
    """Split time-major rollout arrays [T,E,A,...] (numpy or cpu tensors) into rllab-style paths:
    one dict per (env, agent, episode) with ``observations/actions/rewards/env_infos`` arrays, the
    format ``dec_rollout`` returns (ma_sampler.py:88-100).  ``obs[t]`` is the observation the action
    ``actions[t]`` was taken in, i.e. callers pass obs shifted by one step (reset obs first)."""
    obs, actions, rew, done = [np.asarray(x) for x in (obs, actions, rew, done)]
    T, E, A = rew.shape
    infos = {k: np.asarray(v) for k, v in (infos or {}).items()}
    paths = []
    for e in range(E):
        ends = list(np.nonzero(done[:, e])[0] + 1)
        if not ends or ends[-1] != T:
            ends.append(T)
        start = 0
        for end in ends:
            for a in range(A):
                paths.append(dict(observations=obs[start:end, e, a], actions=actions[start:end, e, a],
                                  rewards=rew[start:end, e, a],
                                  env_infos={k: v[start:end, e] for k, v in infos.items()},
                                  env=e, agent=a, terminated=bool(done[end - 1, e])))
            start = end
    return paths
