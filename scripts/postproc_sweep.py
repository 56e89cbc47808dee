"""Achieved HBM bandwidth of the trajectory post-processing passes at the C2 rollout shape
(T = 256, E = 4096, A = 5, D = 213): bytes each pass must read + write / its device time (run under gpurun)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madrl_b200 import BatchedMAWaterWorld  # noqa: E402
from madrl_b200.postproc import (EpisodeStats, FrameStack, Standardizer, center_advantages, explained_variance, gae,  # noqa: E402
                                 pack_paths)

PEAK = 6576.0


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def report(name, nbytes, ms):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print("%-44s %8.3f ms  %7.1f GB/s  (%4.1f%% of %d)  %6.1f MB" % (name, ms, gbs, 100 * gbs / PEAK, PEAK, nbytes / 1e6), flush=True)


def main():
    T, E, A = 256, 4096, 5
    eng = BatchedMAWaterWorld(E, n_pursuers=A, n_evaders=5, seed=0, max_path_length=100)
    D = eng.obs_dim
    obs0 = eng.reset()
    act = torch.randn(T, E, A, 2, device="cuda") * 0.5
    obs, rew, done, info = eng.rollout(act)
    val = torch.randn_like(rew)
    n = T * E * A
    report("gae (rew, value, done -> adv, ret)", 4 * n * 4 + T * E, timeit(lambda: gae(rew, val, done, 0.99, 0.97)))
    adv, ret = gae(rew, val, done, 0.99, 0.97)
    report("center_advantages (2 passes + write)", 3 * n * 4, timeit(lambda: center_advantages(adv)))
    report("explained_variance (2 series, 2 passes)", 4 * n * 4, timeit(lambda: explained_variance(val, ret)))
    st = Standardizer(E, A, D, "cuda", enable_obsnorm=True, enable_rewnorm=True)
    o2 = obs.clone()
    report("Standardizer.obs (in place, f64 mean/var)", 2 * n * D * 4 + 4 * E * A * D * 8, timeit(lambda: st.obs(o2), reps=3))
    r2 = rew.clone()
    report("Standardizer.rew", 2 * n * 4 + 4 * E * A * 8, timeit(lambda: st.rew(r2)))
    es = EpisodeStats(E, A, "cuda")
    report("EpisodeStats.rollout", 2 * n * 4 + 3 * T * E * 4, timeit(lambda: es.rollout(rew, done)))
    Ts = 32   # frame stack x4 of a 32-step slice: output is 4x the input
    fs = FrameStack(E, A, D, 4, "cuda")
    fs.reset(obs0)
    os_, ds_ = obs[:Ts].contiguous(), done[:Ts].contiguous()
    report("FrameStack(4).rollout, T=32", Ts * E * A * D * 4 * (1 + 4), timeit(lambda: fs.rollout(os_, ds_), reps=3))
    infos = dict(evcatches=info[..., 0].contiguous(), pocatches=info[..., 1].contiguous())
    nb = 2 * (n * D * 4 + n * 2 * 4 + n * 4 + 2 * T * E * 4) + T * E
    report("pack_paths (obs, actions, rew, infos -> paths)", nb, timeit(lambda: pack_paths(obs, act, rew, done, infos, obs_before=obs0), reps=3))


if __name__ == "__main__":
    main()
