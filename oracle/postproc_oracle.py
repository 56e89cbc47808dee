"""CPU oracles (TEST INFRASTRUCTURE) for the trajectory post-processing passes: restatements of

* rllab ``BaseSampler.process_samples`` advantage / return computation
  (rllab/rllab/sampler/base.py:48-68; ``discount_cumsum`` = rllab/rllab/misc/special.py:107-111),
* ``ObservationBuffer`` (madrl_environments/__init__.py:143-196),
* ``StandardizedEnv``   (madrl_environments/__init__.py:204-291),
* ``center_advantages`` / ``shift_advantages_to_positive`` (rllab/rllab/algos/util.py:7-12) and
  ``explained_variance_1d`` (rllab/rllab/misc/special.py:51-59),

applied to time-major rollout arrays of ONE env ``[T, A, ...]`` with VecEnvExecutor auto-reset
semantics (the obs slot of a done step already holds the reset observation).
"""
import numpy as np
import scipy.signal


def discount_cumsum(x, discount):                      # special.py:107-111
    return scipy.signal.lfilter([1], [1, float(-discount)], x[::-1], axis=0)[::-1]


def gae_env(rew, val, done, discount, gae_lambda, last_value=None):
    """rew, val [T, A]; done [T] -> adv, ret [T, A] (float64)."""
    T, A = rew.shape
    adv, ret = np.zeros((T, A)), np.zeros((T, A))
    ends = list(np.nonzero(done)[0] + 1)
    tail_open = not ends or ends[-1] != T
    if tail_open:
        ends.append(T)
    start = 0
    for n, end in enumerate(ends):
        boot = np.zeros(A)
        if tail_open and n == len(ends) - 1 and last_value is not None:
            boot = np.asarray(last_value, dtype=np.float64)
        for a in range(A):
            r = rew[start:end, a].astype(np.float64)
            b = np.append(val[start:end, a].astype(np.float64), boot[a])          # base.py:57
            deltas = r + discount * b[1:] - b[:-1]                                # base.py:58-60
            adv[start:end, a] = discount_cumsum(deltas, discount * gae_lambda)    # base.py:61-62
            # returns bootstrap the same way when a tail value is supplied
            ret[start:end, a] = discount_cumsum(np.append(r, boot[a]), discount)[:-1] \
                if boot[a] != 0 else discount_cumsum(r, discount)                 # base.py:63
        start = end
    return adv, ret


def frame_stack_env(obs0, obs, done, B):
    """obs0 [A, D] reset obs; obs [T, A, D]; done [T] -> (stack0 [A,D,B], stacked [T,A,D,B])."""
    A, D = obs0.shape
    buf = np.zeros((A, D, B))
    for b in range(B):                                  # reset(): __init__.py:186-196
        buf[..., b] = obs0
    out0 = buf.copy()
    out = np.zeros((obs.shape[0], A, D, B))
    for t in range(obs.shape[0]):
        if done[t]:                                     # executor called wrapper.reset()
            for b in range(B):
                buf[..., b] = obs[t]
        else:                                           # step(): __init__.py:176-183
            buf[..., 0:B - 1] = buf[..., 1:B].copy()
            buf[..., -1] = obs[t]
        out[t] = buf
    return out0, out


class StandardizeEnv(object):
    """Running estimates of one env (one per agent, __init__.py:221-234)."""

    def __init__(self, A, D, scale_reward=1., enable_obsnorm=False, enable_rewnorm=False,
                 obs_alpha=0.001, rew_alpha=0.001, eps=1e-8):
        self.scale_reward, self.eo, self.er = scale_reward, enable_obsnorm, enable_rewnorm
        self.oa, self.ra, self.eps = obs_alpha, rew_alpha, eps
        self.obs_mean, self.obs_var = np.zeros((A, D)), np.ones((A, D))
        self.rew_mean, self.rew_var = np.zeros(A), np.ones(A)

    def obs(self, x):                                   # standardize_obs, __init__.py:258-262
        if not self.eo:
            return x
        self.obs_mean = (1 - self.oa) * self.obs_mean + self.oa * x
        self.obs_var = (1 - self.oa) * self.obs_var + self.oa * np.square(x - self.obs_mean)
        return (x - self.obs_mean) / (np.sqrt(self.obs_var) + self.eps)

    def rew(self, r):                                   # standardize_rew + scale, __init__.py:264-289
        if self.er:
            self.rew_mean = (1 - self.ra) * self.rew_mean + self.ra * r
            self.rew_var = (1 - self.ra) * self.rew_var + self.ra * np.square(r - self.rew_mean)
            r = r / (np.sqrt(self.rew_var) + self.eps)
        return self.scale_reward * r


def episode_stats_env(rew, done, discount=0.99, max_traj_len=500):
    """DiagnosticsWrapper (madrl_environments/__init__.py:314-369) on one env's rollout arrays
    rew [T, A], done [T]: list of (t, episode_reward [A], avg, disc_return, length) records."""
    A = rew.shape[1]
    ep_reward, length, all_rewards, out = np.zeros(A), 0, [], []
    for t in range(rew.shape[0]):
        ep_reward += np.asarray(rew[t], dtype=np.float64)
        length += 1
        all_rewards.append(np.asarray(rew[t], dtype=np.float64))
        if done[t] or length >= max_traj_len:                      # __init__.py:352
            arr = np.asarray(all_rewards).mean(axis=1)
            disc = np.sum(arr * (discount ** np.arange(len(arr))))   # _discount_sum, __init__.py:392
            out.append((t, ep_reward.copy(), float(np.mean(ep_reward)), float(disc), length))
            ep_reward, length, all_rewards = np.zeros(A), 0, []
    return out


def center_advantages(adv):                             # rllab/rllab/algos/util.py:7-8
    adv = np.asarray(adv, dtype=np.float64)
    return (adv - np.mean(adv)) / (adv.std() + 1e-8)


def shift_advantages_to_positive(adv):                  # rllab/rllab/algos/util.py:11-12
    adv = np.asarray(adv, dtype=np.float64)
    return (adv - np.min(adv)) + 1e-8


def explained_variance_1d(ypred, y):                    # rllab/rllab/misc/special.py:51-59
    ypred, y = np.asarray(ypred, dtype=np.float64), np.asarray(y, dtype=np.float64)
    assert y.ndim == 1 and ypred.ndim == 1
    vary = np.var(y)
    if np.isclose(vary, 0):
        return 0 if np.var(ypred) > 0 else 1
    return 1 - np.var(y - ypred) / (vary + 1e-8)
