"""Throughput of the ContinuousHostageWorld rollout kernel (run under gpurun)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from madrl_b200 import BatchedHostageWorld


def bytes_per_env_step(Nr, Nh, Nc, K):
    # SURVEY.md 8(d)
    return 4 * (2 * 4 * Nr + 2 * 4 * Nc + 2 * Nh + 4) + 2 * ((Nh + 7) // 8 + 12 + 8) + 4 * 2 * Nr + \
        4 * Nr * (5 * K + 6) + 4 * Nr + 9


def run(E, T, args=(10, 16, 16, 4, 2), reps=5, mpl=0):
    eng = BatchedHostageWorld(E, *args, seed=0, max_path_length=mpl)
    eng.reset()
    Nr = args[0]
    act = torch.randn(T, E, Nr, 2, device='cuda') * 0.5
    out = (torch.empty((T, E, Nr, eng.obs_dim), device='cuda'), torch.empty((T, E, Nr), device='cuda'),
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
    bpe = bytes_per_env_step(Nr, args[1], args[2], 30)
    gbs = bpe * E * T / (ms * 1e-3) / 1e9
    print("hostage E=%d T=%d: %.3f ms/launch  %.2f us/step  %.1f GB/s (%.1f%% of 6576)  %.2f G agent-steps/s"
          % (E, T, ms, 1e3 * ms / T, gbs, 100 * gbs / 6576, E * Nr * T / ms / 1e6), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        run(8192, 32)
        sys.exit(0)
    run(8192, 32)
    run(4096, 64)
    run(65536, 8)
