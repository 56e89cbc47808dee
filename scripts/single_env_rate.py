"""Steps/s of the single-env drop-in classes (the compatibility path of INTEGRATION.md section 1): one kernel
launch + host round trips per step().  Run under gpurun; prints one line per family."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madrl_b200 import ContinuousHostageWorld, MAWaterWorld, PursuitEvade  # noqa: E402


def rate(env, act, n=300):
    env.reset()
    for _ in range(20):
        env.step(act())
    t0 = time.perf_counter()
    for _ in range(n):
        _, _, done, _ = env.step(act())
        if done:
            env.reset()
    return n / (time.perf_counter() - t0)


if __name__ == "__main__":
    rs = np.random.RandomState(0)
    ww = MAWaterWorld(5, 5)
    print("MAWaterWorld(5, 5).step(): %.0f steps/s" % rate(ww, lambda: rs.randn(5, 2) * 0.5))
    maps = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "maps", "map_pool16.npy"))
    pe = PursuitEvade(maps, n_evaders=30, n_pursuers=8, obs_range=7, surround=True, n_catch=2, flatten=True,
                      reward_mech='local', catchr=0.1, term_pursuit=5.0, sample_maps=True)
    print("PursuitEvade(map_pool16, 8p/30e).step(): %.0f steps/s" % rate(pe, lambda: list(rs.randint(0, 5, 8))))
    hw = ContinuousHostageWorld(10, 16, 16, 4, 2)
    print("ContinuousHostageWorld(10, 16, 16).step(): %.0f steps/s" % rate(hw, lambda: rs.randn(10, 2) * 0.5))
