"""Host-side mirror of the reference's multi-agent env interface.

``Agent`` / ``AbstractMAEnv`` follow ``madrl_environments/__init__.py:9-119`` (same method and
property names, same ``set_param_values`` contract); ``EzPickle`` follows
``rltools/rltools/util.py:261-288`` (pickle = re-run the constructor with the saved arguments),
which is what rllab relies on when it ships an env to sampler workers
(``rllab/rllab/sampler/ma_sampler.py:182-197``).
"""
import numpy as np


class EzPickle(object):
    def __init__(self, *args, **kwargs):
        self._ezpickle_args = args
        self._ezpickle_kwargs = kwargs

    def __getstate__(self):
        return {"_ezpickle_args": self._ezpickle_args, "_ezpickle_kwargs": self._ezpickle_kwargs}

    def __setstate__(self, d):
        out = type(self)(*d["_ezpickle_args"], **d["_ezpickle_kwargs"])
        self.__dict__.update(out.__dict__)


class Agent(object):
    @property
    def observation_space(self):
        raise NotImplementedError()

    @property
    def action_space(self):
        raise NotImplementedError()

    def __str__(self):
        return '<{} instance>'.format(type(self).__name__)


class AbstractMAEnv(object):
    """Same surface as madrl_environments.AbstractMAEnv (``__init__.py:27-119``)."""

    _unwrapped = None

    def setup(self):
        pass

    def seed(self, seed=None):
        return []

    @property
    def agents(self):
        raise NotImplementedError()

    @property
    def reward_mech(self):
        raise NotImplementedError()

    def reset(self):
        raise NotImplementedError()

    def step(self, actions):
        raise NotImplementedError()

    @property
    def is_terminal(self):
        raise NotImplementedError()

    def set_param_values(self, lut):
        for k, v in lut.items():
            setattr(self, k, v)
        self.setup()

    def render(self, *args, **kwargs):
        raise NotImplementedError("rendering is outside the accelerated hot path")

    def animate(self, act_fn, nsteps, **kwargs):
        """Roll the env with per-agent action functions (``__init__.py:72-109``, minus video)."""
        if not isinstance(act_fn, list):
            act_fn = [act_fn for _ in range(len(self.agents))]
        assert len(act_fn) == len(self.agents)
        obs = self.reset()
        rew = np.zeros((len(self.agents)))
        infos = []
        for _ in range(nsteps):
            a = list(map(lambda afn, o: afn(o), act_fn, obs))
            obs, r, done, info = self.step(a)
            rew += r
            if info:
                infos.append(info)
            if done:
                break
        keys = infos[0].keys() if infos else []
        return rew, {k: np.array([i[k] for i in infos]) for k in keys}

    @property
    def unwrapped(self):
        return self._unwrapped if self._unwrapped is not None else self

    def __str__(self):
        return '<{} instance>'.format(type(self).__name__)
