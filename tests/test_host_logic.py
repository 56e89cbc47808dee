"""Host-side logic that needs no GPU: space descriptors, the env-interface mirror, config
validation through the C ABI's host-only layout functions, build staleness."""
import ctypes as C
import pickle

import numpy as np
import pytest

from madrl_b200 import _lib
from madrl_b200.core import AbstractMAEnv, Agent, EzPickle
from madrl_b200.spaces import Box, Discrete


def test_spaces_follow_gym_construction_rules():
    b = Box(low=-10, high=10, shape=(213,))
    assert b.shape == (213,) and b.low.min() == -10 and b.high.max() == 10
    assert b.contains(np.zeros(213)) and not b.contains(np.full(213, 11.0)) and not b.contains(np.zeros(3))
    b2 = Box(np.zeros(148), np.ones(148))
    assert b2.shape == (148,) and b2 == Box(np.zeros(148), np.ones(148))
    assert b.sample().shape == (213,)
    d = Discrete(5)
    assert d.n == 5 and d.contains(4) and not d.contains(5) and 0 <= d.sample() < 5 and d == Discrete(5)


class _Toy(AbstractMAEnv, EzPickle):
    class _A(Agent):
        observation_space = Box(low=0, high=1, shape=(2,))
        action_space = Discrete(2)

    def __init__(self, n, scale=1.0):
        EzPickle.__init__(self, n, scale=scale)
        self.n, self.scale, self.t, self.setups = n, scale, 0, 0

    def setup(self):
        self.setups += 1

    @property
    def agents(self):
        return [self._A() for _ in range(self.n)]

    @property
    def reward_mech(self):
        return 'local'

    def reset(self):
        self.t = 0
        return [np.zeros(2) for _ in range(self.n)]

    def step(self, a):
        self.t += 1
        return [np.full(2, self.t) for _ in range(self.n)], np.full(self.n, self.scale), self.t >= 3, {'k': self.t}


def test_env_interface_mirror():
    env = _Toy(2, scale=0.5)
    assert env.unwrapped is env and str(env) == '<_Toy instance>' and str(env.agents[0]) == '<_A instance>'
    env.set_param_values(dict(scale=2.0))                  # setattr + setup(), __init__.py:64-67
    assert env.scale == 2.0 and env.setups == 1
    rew, info = env.animate(lambda o: 0, 10)               # stops at done, stacks infos
    assert list(rew) == [6.0, 6.0] and list(info['k']) == [1, 2, 3]
    clone = pickle.loads(pickle.dumps(_Toy(3, scale=0.25)))  # EzPickle: rebuilt from ctor args
    assert clone.n == 3 and clone.scale == 0.25 and clone.t == 0
    with pytest.raises(NotImplementedError):
        env.render()


def test_layout_validation_errors_are_reported():
    lib = _lib.lib()
    pe = _lib.PEConfig(n_envs=4, n_pursuers=8, n_evaders=70, xs=16, ys=16, n_maps=1, obs_range=7, flatten=1,
                       layer_norm=10, constraint_window=1.0)
    lay = _lib.PELayout()
    assert lib.madrl_pursuit_state_layout(C.byref(pe), C.byref(lay)) == -1 and b"n_evaders" in lib.madrl_last_error()
    pe.n_evaders = 30
    pe.constraint_window = 0.0
    assert lib.madrl_pursuit_state_layout(C.byref(pe), C.byref(lay)) == -1 and b"constraint_window" in lib.madrl_last_error()
    pe.constraint_window = 1.0
    assert lib.madrl_pursuit_state_layout(C.byref(pe), C.byref(lay)) == 0 and lay.obs_dim == 147 and lay.n_agents == 38
    pe.flatten, pe.include_id = 0, 1
    assert lib.madrl_pursuit_state_layout(C.byref(pe), C.byref(lay)) == 0 and lay.obs_dim == 4 * 49
    hw = _lib.HWConfig(n_envs=2, n_good=10, n_hostages=16, n_bad=16, n_coop_save=4, n_coop_avoid=2, n_sensors=30,
                       addid=1, timestep_limit=1000)
    hl = _lib.HWLayout()
    assert lib.madrl_hostage_state_layout(C.byref(hw), C.byref(hl)) == 0 and hl.obs_dim == 156 and hl.n_obj == 42
    hw.n_good = 40
    assert lib.madrl_hostage_state_layout(C.byref(hw), C.byref(hl)) == -1 and b"n_good" in lib.madrl_last_error()
    assert lib.madrl_gae_f32(0, 1, 1, None, None, None, None, 0.99, 0.95, None, None, None) == -1


def test_engines_refuse_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from madrl_b200 import BatchedMAWaterWorld, EngineError
    with pytest.raises(EngineError, match="no CPU fallback"):
        BatchedMAWaterWorld(4, 5, 5)


def test_shard_construction_arguments():
    from madrl_b200.dist import make_sharded, shard_range
    made = {}

    class Fake(object):
        def __init__(self, n, *args, env_id_base=0, **kw):
            made.update(n=n, base=env_id_base, args=args, kw=kw)
    eng = make_sharded(Fake, 4097, 5, 5, rank=3, world=8, seed=1)
    lo, hi = shard_range(4097, 3, 8)
    assert made == dict(n=hi - lo, base=lo, args=(5, 5), kw=dict(seed=1)) and eng.shard == (lo, hi, 4097)
    with pytest.raises(ValueError):
        make_sharded(Fake, 2, rank=0, world=8)      # rank 0 of 8 owns nothing of a 2-env batch


def test_require_tensor_rejects_wrong_dtype_shape_layout_and_device():
    """caller-supplied buffers are validated before their data_ptr() crosses the C ABI."""
    import pytest
    import torch
    from madrl_b200._lib import require_tensor
    ok = torch.zeros((3, 4, 5), dtype=torch.float32)
    assert require_tensor(ok, "obs", torch.float32, (3, 4, 5), 'cpu') is ok
    with pytest.raises(TypeError):
        require_tensor(ok.double(), "obs", torch.float32, (3, 4, 5), 'cpu')          # fp64 on an fp32 engine
    with pytest.raises(ValueError):
        require_tensor(ok[:2], "obs", torch.float32, (3, 4, 5), 'cpu')               # short buffer
    with pytest.raises(ValueError):
        require_tensor(ok.transpose(0, 1), "obs", torch.float32, (4, 3, 5), 'cpu')   # non-contiguous view
    with pytest.raises(ValueError):
        require_tensor(ok, "obs", torch.float32, (3, 4, 5), torch.device("cuda", 0))  # host tensor for a device slot
    with pytest.raises(TypeError):
        require_tensor(ok.numpy(), "obs", torch.float32, (3, 4, 5), 'cpu')


def test_stale_library_is_detected_by_content_hash(tmp_path, monkeypatch):
    """a git-ignored .so built from other sources must not be loaded silently."""
    from madrl_b200 import build as B
    assert not B._stale()                     # conftest / an earlier test built it from this tree
    monkeypatch.setattr(B, "NVCC_FLAGS", B.NVCC_FLAGS + ["-DSOMETHING_ELSE"])
    assert B._stale()                         # same files, other flags: another binary
