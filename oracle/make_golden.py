"""Generate golden input/output vectors from the REAL reference (run in the build container).

    python -m oracle.make_golden

Drives the unmodified reference classes (via ``oracle/refshim.py``) with the counter-based
streams of ``oracle/philox.py`` injected where the reference draws random numbers, and writes
small ``.npz`` fixtures to ``tests/golden/``.  The reference tree does not exist on the GPU box;
these files (and this script, for provenance) are what travels.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle.philox import Stream  # noqa: E402
from oracle.refshim import load_reference  # noqa: E402

WW_CASES = {
    # name: (ctor kwargs, seed, env_id, T, action_std)
    "ww_c2": (dict(n_pursuers=5, n_evaders=5), 11, 3, 120, 0.5),
    "ww_dense": (dict(n_pursuers=5, n_evaders=5, n_coop=1, radius=0.04, sensor_range=0.3), 12, 0, 120, 1.0),
    "ww_c4": (dict(n_pursuers=20, n_evaders=50, n_poison=50), 13, 40000, 16, 0.5),
    "ww_global_nospeed": (dict(n_pursuers=3, n_evaders=4, n_poison=2, n_sensors=7, n_coop=1, radius=0.05,
                               reward_mech='global', speed_features=False, addid=False,
                               obstacle_loc=None), 14, 9, 150, 1.0),
}


def gen_waterworld(MAWaterWorld):
    for name, (kw, seed, env_id, T, std) in WW_CASES.items():
        env = MAWaterWorld(**kw)
        env.np_random = Stream(seed, env_id)
        Np = kw['n_pursuers']
        obs0 = np.array(env.reset())
        arng = np.random.RandomState(seed)
        actions = (arng.randn(T, Np, 2) * std).astype(np.float32).astype(np.float64)
        obs, rew, done, info = [], [], [], []
        for t in range(T):
            o, r, d, i = env.step(actions[t])
            obs.append(np.array(o)); rew.append(np.array(r)); done.append(d)
            info.append([i['evcatches'], i['pocatches']])
        cfg = {k: (None if v is None else v) for k, v in kw.items()}
        np.savez_compressed(
            os.path.join(GOLDEN, name + ".npz"), config=json.dumps(cfg), seed=seed, env_id=env_id,
            actions=actions, obs0=obs0, obs=np.array(obs), rew=np.array(rew),
            done=np.array(done), info=np.array(info, dtype=np.int32),
            final_px=np.array([p.position for p in env._pursuers]),
            final_ex=np.array([p.position for p in env._evaders]),
            final_ov=np.array([p.velocity for p in env._poisons]),
            counter=env.np_random.counter)
        print(name, "catches", np.array(info).sum(0), "draws", env.np_random.counter)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    MAWaterWorld, PursuitEvade, ContinuousHostageWorld = load_reference()
    gen_waterworld(MAWaterWorld)


if __name__ == "__main__":
    main()
