"""Throughput of the PursuitEvade rollout kernel (run under gpurun)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from madrl_b200 import BatchedPursuitEvade

C3 = dict(n_evaders=30, n_pursuers=8, obs_range=7, surround=True, n_catch=2, flatten=True,
          reward_mech='local', catchr=0.1, term_pursuit=5.0, sample_maps=True, include_id=True)


def bytes_per_env_step(Np, Ne, R):
    # state r+w, actions, obs, rewards, done, info
    return 2 * (2 * Np + 2 * Ne + (Ne + 7) // 8 + 1 + 8) + 4 * Np + 4 * Np * (3 * R * R + 1) + 4 * Np + 1 + 4


def run(E, T, wpb=0, bps=0, cfg=C3, reps=5, mpl=500):
    maps = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "maps", "map_pool16.npy"))
    eng = BatchedPursuitEvade(E, maps, seed=0, max_path_length=mpl, **cfg)
    eng.set_launch(wpb, bps)
    eng.reset()
    Np = cfg['n_pursuers']
    act = torch.randint(0, 5, (T, E, Np), dtype=torch.int32, device='cuda')
    out = (torch.empty((T, E, Np, eng.obs_dim), device='cuda'), torch.empty((T, E, Np), device='cuda'),
           torch.empty((T, E), dtype=torch.uint8, device='cuda'), torch.empty((T, E), dtype=torch.int32, device='cuda'))
    for _ in range(2):
        eng.rollout(act, out=out)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        eng.rollout(act, out=out)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    bpe = bytes_per_env_step(Np, cfg['n_evaders'], cfg['obs_range'])
    gbs = bpe * E * T / (ms * 1e-3) / 1e9
    print("pursuit E=%d T=%d wpb=%d bps=%d: %.3f ms/launch  %.2f us/step  %.1f GB/s (%.1f%% of 6576)  %.2f G agent-steps/s"
          % (E, T, wpb, bps, ms, 1e3 * ms / T, gbs, 100 * gbs / 6576, E * Np * T / ms / 1e6), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        run(65536, 32)
        sys.exit(0)
    for wpb in (4, 2):
        run(65536, 8, wpb)
    run(65536, 32)
    run(16384, 32)
    run(4096, 64)
    run(65536, 1, reps=10)
