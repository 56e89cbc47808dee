"""Open-loop vs closed-loop (in-kernel heuristic policy) rollout kernel times (run under gpurun)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madrl_b200 import BatchedMAWaterWorld, BatchedPursuitEvade  # noqa: E402


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


def main():
    for name, cfg, E, T in (("c2", dict(n_pursuers=5, n_evaders=5), 4096, 256),
                            ("c4", dict(n_pursuers=20, n_evaders=50, n_poison=50), 4096, 64)):
        eng = BatchedMAWaterWorld(E, seed=0, **cfg)
        obs0 = eng.reset()
        Np = cfg["n_pursuers"]
        out = (torch.empty((T, E, Np, eng.obs_dim), device="cuda"), torch.empty((T, E, Np), device="cuda"),
               torch.empty((T, E), dtype=torch.uint8, device="cuda"), torch.empty((T, E, 2), dtype=torch.int32, device="cuda"))
        act = torch.randn(T, E, Np, 2, device="cuda") * 0.5
        t_open = timeit(lambda: eng.rollout(act, out=out))
        state = {"o": obs0}

        def closed(rec):
            r = eng.rollout_heuristic(T, state["o"], out=out, record_actions=rec)
            state["o"] = r[1][-1]
        t_rec = timeit(lambda: closed(True))
        t_norec = timeit(lambda: closed(False))
        print("waterworld %s E=%d T=%d: open loop %.3f ms, closed loop %.3f ms (actions recorded) / %.3f ms (not recorded)"
              % (name, E, T, t_open, t_rec, t_norec), flush=True)
        # same dynamics both ways: replay the actions the policy took, from the same state, through the open-loop kernel
        e1 = BatchedMAWaterWorld(E, seed=0, **cfg)
        o1 = e1.reset()
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        recs, prev = [], o1
        torch.cuda.synchronize()
        s_.record()
        for _ in range(4):
            r = e1.rollout_heuristic(T, prev, out=out)
            recs.append(r[0]); prev = r[1][-1].clone()
        e_.record()
        torch.cuda.synchronize()
        t_cl = s_.elapsed_time(e_) / 4
        e2 = BatchedMAWaterWorld(E, seed=0, **cfg)
        e2.reset()
        torch.cuda.synchronize()
        s_.record()
        for a_ in recs:
            e2.rollout(a_, out=out)
        e_.record()
        torch.cuda.synchronize()
        print("   same trajectory (first 4 rollouts after reset): closed loop %.3f ms, open-loop replay of its actions %.3f ms"
              % (t_cl, s_.elapsed_time(e_) / 4), flush=True)
    maps = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "maps", "map_pool16.npy"))
    cfg = dict(n_evaders=30, n_pursuers=8, obs_range=7, surround=True, n_catch=2, flatten=True, reward_mech='local',
               catchr=0.1, term_pursuit=5.0, sample_maps=True, include_id=True)
    E, T = 65536, 16
    pe = BatchedPursuitEvade(E, maps, seed=0, max_path_length=500, **cfg)
    obs0 = pe.reset()
    out = (torch.empty((T, E, 8, pe.obs_dim), device="cuda"), torch.empty((T, E, 8), device="cuda"),
           torch.empty((T, E), dtype=torch.uint8, device="cuda"), torch.empty((T, E), dtype=torch.int32, device="cuda"))
    act = torch.randint(0, 5, (T, E, 8), dtype=torch.int32, device="cuda")
    t_open = timeit(lambda: pe.rollout(act, out=out))
    state = {"o": obs0}

    def closed():
        r = pe.rollout_heuristic(T, state["o"], out=out)
        state["o"] = r[1][-1]
    t_closed = timeit(closed)
    print("pursuit c3 E=%d T=%d: open loop %.3f ms, closed loop %.3f ms" % (E, T, t_open, t_closed), flush=True)


if __name__ == "__main__":
    main()
