"""Launch-geometry sweep of the Waterworld rollout kernel (run under gpurun)."""
import itertools
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from madrl_b200 import BatchedMAWaterWorld

def run(E, T, wpb, bps, cfg, reps=5):
    eng = BatchedMAWaterWorld(E, seed=0, **cfg)
    eng.set_launch(wpb, bps)
    eng.reset()
    Np = cfg['n_pursuers']
    act = torch.randn(T, E, Np, 2, device='cuda') * 0.5
    out = (torch.empty((T, E, Np, eng.obs_dim), device='cuda'), torch.empty((T, E, Np), device='cuda'),
           torch.empty((T, E), dtype=torch.uint8, device='cuda'), torch.empty((T, E, 2), dtype=torch.int32, device='cuda'))
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
    Ne, Npo, K = cfg['n_evaders'], cfg.get('n_poison', 10), cfg.get('n_sensors', 30)
    bpe = 4 * (8 * (Np + Ne + Npo) + 2 * Np + Np * (7 * K + 3) + Np) + 41
    gbs = bpe * E * T / (ms * 1e-3) / 1e9
    print("E=%d T=%d wpb=%d bps=%d: %.3f ms/launch  %.2f us/step  %.1f GB/s (%.1f%% of 6576)  %.2f G agent-steps/s"
          % (E, T, wpb, bps, ms, 1e3 * ms / T, gbs, 100 * gbs / 6576, E * Np * T / ms / 1e6), flush=True)

if __name__ == "__main__":
    c2 = dict(n_pursuers=5, n_evaders=5)
    c4 = dict(n_pursuers=20, n_evaders=50, n_poison=50)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "c2"):
        for wpb, bps in [(4, 0), (2, 0)]:
            run(4096, 64, wpb, bps, c2)
        for T in (1, 8, 256):
            run(4096, T, 4, 0, c2, reps=20 if T == 1 else 5)
        for E in (2048, 8192, 16384, 65536):
            run(E, 16, 4, 0, c2)
    if which == "c2one":
        run(4096, 256, 4, 0, c2, reps=2)
    if which == "c4one":
        run(4096, 16, 4, 0, c4, reps=2)
    if which == "quick":
        run(4096, 256, 4, 0, c2)
        run(4096, 256, 4, 0, c2)
        run(4096, 16, 4, 0, c4)
        run(4096, 64, 4, 0, c4)
    if which in ("all", "c4"):
        run(4096, 16, 4, 0, c4)
        run(16384, 8, 4, 0, c4)
