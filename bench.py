#!/usr/bin/env python
"""Headline benchmark: agent-env-steps/s of the MAWaterWorld rollout hot path (BASELINE.json
configs[1]: 5 pursuers / 5 evaders / 10 poison / 30 sensors, 4096 envs per B200).

    python bench.py [--gpus N --steps K --warmup W]            # this engine
    python bench.py --impl reference [...]                      # the CPU path on host cores
    torchrun --nproc-per-node N bench.py --gpus N ...           # one rank per GPU, weak scaling

One bench "step" = one rollout launch = T_INNER lockstep env steps of every env in the batch
(one pass of the hot path over one batch of synthetic actions).  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WW_CFG = dict(n_pursuers=5, n_evaders=5, n_poison=10, n_sensors=30)   # class defaults otherwise
METRIC = "agent_env_steps_per_sec"
UNIT = "agent-env-steps/s"

PE_C3 = dict(n_evaders=30, n_pursuers=8, obs_range=7, surround=True, n_catch=2, flatten=True,
             reward_mech='local', catchr=0.1, term_pursuit=5.0, sample_maps=True, include_id=True)
HW_C5 = (10, 16, 16, 4, 2)

# name -> description of the BASELINE.json config it measures (default = configs[1])
WORKLOADS = {
    "waterworld": dict(desc="MAWaterWorld 5p/5e/10po/30 sensors", envs=4096, agents=5, family="ww", cfg=WW_CFG),
    "waterworld_c4": dict(desc="MAWaterWorld 20p/50e/50po/30 sensors", envs=4096, agents=20, family="ww",
                          cfg=dict(n_pursuers=20, n_evaders=50, n_poison=50, n_sensors=30)),
    "pursuit": dict(desc="PursuitEvade 16x16 map_pool16, 8p/30e, obs_range 7, surround", envs=65536, agents=8,
                    family="pe", cfg=PE_C3),
    "hostage": dict(desc="ContinuousHostageWorld 10 rescuers/16 hostages/16 criminals/30 sensors", envs=8192,
                    agents=10, family="hw", cfg=HW_C5),
}


def bytes_per_env_step(wl):
    """Algorithmic (compulsory) bytes per env-step, SURVEY.md 8(d)."""
    w = WORKLOADS[wl]
    if w["family"] == "ww":
        c = w["cfg"]
        return ww_bytes_per_env_step(c["n_pursuers"], c["n_evaders"], c["n_poison"], c["n_sensors"])
    if w["family"] == "pe":
        Np, Ne, R = 8, 30, 7
        return 2 * (2 * Np + 2 * Ne + (Ne + 7) // 8 + 1 + 8) + 4 * Np + 4 * Np * (3 * R * R + 1) + 4 * Np + 1 + 4
    Nr, Nh, Nc, K = 10, 16, 16, 30
    return 4 * (2 * 4 * Nr + 2 * 4 * Nc + 2 * Nh + 4) + 2 * ((Nh + 7) // 8 + 12 + 8) + 4 * 2 * Nr + \
        4 * Nr * (5 * K + 6) + 4 * Nr + 9


def make_engine(wl, E, dev, rank, mpl=0):
    from madrl_b200 import BatchedHostageWorld, BatchedMAWaterWorld, BatchedPursuitEvade
    w = WORKLOADS[wl]
    if w["family"] == "ww":
        return BatchedMAWaterWorld(E, device=dev, seed=0, env_id_base=rank * E, **w["cfg"])
    if w["family"] == "pe":
        maps = np.load(os.path.join(ROOT, "maps", "map_pool16.npy"))
        return BatchedPursuitEvade(E, maps, device=dev, seed=0, env_id_base=rank * E, max_path_length=500, **w["cfg"])
    return BatchedHostageWorld(E, *w["cfg"], device=dev, seed=0, env_id_base=rank * E)


def make_actions(wl, T, E, dev, g, host=False):
    import torch
    w = WORKLOADS[wl]
    d = "cpu" if host else dev
    gg = None if host else g
    if w["family"] == "pe":
        return torch.randint(0, 5, (T, E, w["agents"]), dtype=torch.int32, device=d, generator=gg)
    return torch.randn(T, E, w["agents"], 2, device=d, generator=gg) * 0.5


def make_oracle(wl, seed):
    from oracle.philox import Stream
    w = WORKLOADS[wl]
    if w["family"] == "ww":
        from oracle.waterworld_oracle import WaterworldOracle
        return WaterworldOracle(rng=Stream(seed, seed), **w["cfg"])
    if w["family"] == "pe":
        from oracle.pursuit_oracle import PursuitOracle
        maps = np.load(os.path.join(ROOT, "maps", "map_pool16.npy"))
        return PursuitOracle(maps, rng=Stream(seed, seed), **w["cfg"])
    from oracle.hostage_oracle import HostageOracle
    return HostageOracle(*w["cfg"], rng=Stream(seed, seed))


def ww_bytes_per_env_step(Np, Ne, Npo, K):
    """Algorithmic (compulsory) bytes per env-step, SURVEY.md 8(d) / BASELINE.md section 5."""
    return 4 * (8 * (Np + Ne + Npo) + 2 * Np + Np * (7 * K + 3) + Np) + 41


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace('.', '').isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


# ----------------------------------------------------------------------------- CPU baseline
def _cpu_worker(args):
    seconds, seed, wl = args
    w = WORKLOADS[wl]
    env = make_oracle(wl, seed)
    env.reset()
    rs = np.random.RandomState(seed)
    if w["family"] == "pe":
        acts = rs.randint(0, 5, size=(4096, w["agents"]))
    else:
        acts = rs.randn(4096, w["agents"] * 2) * 0.5
    for i in range(50):
        env.step(acts[i])
    n, last_reset, t0 = 0, 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        _, _, done, _ = env.step(acts[n % 4096])
        n += 1
        if done or n - last_reset >= 500:
            env.reset()
            last_reset = n
    return n, time.perf_counter() - t0


def cpu_baseline(seconds, procs, wl="waterworld"):
    """The CPU path (float64 NumPy oracle port of the reference step(), one env per process the
    way rllab's StatefulPool parallelises; rllab/rllab/sampler/stateful_pool.py:102-157)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_worker, [(seconds, 100 + i, wl) for i in range(procs)])
    env_steps_per_s = sum(n / dt for n, dt in res)
    return env_steps_per_s * WORKLOADS[wl]["agents"]


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="waterworld", choices=sorted(WORKLOADS),
                    help="default = BASELINE.json configs[1]; the others are the remaining configs")
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (weak scaling); 0 = the config's")
    ap.add_argument("--t-inner", type=int, default=0, help="lockstep env steps per rollout launch (0 = auto)")
    ap.add_argument("--cpu-seconds", type=float, default=8.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--gather-obs", action="store_true",
                    help="also all-gather the full obs tensor every rollout (NVLink-bound)")
    ap.add_argument("--exchange", default="overlap", choices=["overlap", "nccl", "fused", "fused-all"],
                    help="multi-GPU per-rollout trajectory exchange (see the comment in main())")
    ap.add_argument("--wpb", type=int, default=0)
    ap.add_argument("--bps", type=int, default=0)
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    WL = WORKLOADS[a.workload]
    Np = WL["agents"]
    if a.envs <= 0:
        a.envs = WL["envs"]
    if a.t_inner <= 0:      # keep one launch's obs output around 4.5 GB
        a.t_inner = 256 if a.workload == "waterworld" else (16 if a.workload == "pursuit" else 64)
    host_cores = os.cpu_count() or 1

    if a.impl == "reference":
        if rank != 0:
            return
        procs = host_cores
        per = max(2.0, min(10.0, 60.0 / max(1, a.steps + a.warmup)))
        for _ in range(a.warmup):
            cpu_baseline(0.5, procs, a.workload)
        t0 = time.perf_counter()
        vals = [cpu_baseline(per, procs, a.workload) for _ in range(a.steps)]
        dt = time.perf_counter() - t0
        v = float(np.mean(vals))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WL["desc"] + ", one env per host process", "envs_per_gpu": a.envs},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": procs, "kind": "port",
                             "sample": "%d processes x %.1f s of step() per bench step, float64 NumPy "
                                       "oracle port of the reference step() (the reference tree cannot "
                                       "travel to the GPU box)" % (procs, per)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    import torch
    import torch.distributed as dist
    from madrl_b200 import launch_count

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    E, T = a.envs, a.t_inner
    eng = make_engine(a.workload, E, dev, rank)
    if a.wpb or a.bps:
        eng.set_launch(a.wpb, a.bps)
    D = eng.obs_dim
    eng.reset()
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    # synthetic actions 0.5*N(0,1) (waterworld.py:486), a different tensor per timed step
    n_act = min(a.steps, 4)
    actions = [make_actions(a.workload, T, E, dev, g) for _ in range(n_act)]
    from madrl_b200.dist import PackedTrajectory
    info_w = 1 if WL["family"] == "pe" else 2
    # Multi-GPU exchange of the per-rollout trajectory tensors (rewards / dones / infos; obs stays
    # sharded with the data-parallel learner unless --gather-obs):
    #   overlap (default) one copy-engine P2P copy of the packed buffer to rank 0 over NVLink on a side
    #                     stream + a completion all-reduce, overlapping the next rollout
    #                     (madrl_b200.dist.AsyncRootGather)
    #   nccl              one packed NCCL all_gather after each rollout (serialised with compute)
    #   fused / fused-all the rollout kernel itself stores the rows into rank 0's / every rank's
    #                     buffers over NVLink peer memory (madrl_b200.dist.PeerGather)
    mode, note = ("none" if world == 1 else a.exchange), None
    peer = agather = None
    if world > 1 and mode in ("overlap", "fused", "fused-all"):
        from madrl_b200.dist import AsyncRootGather, PeerGather
        try:   # AsyncRootGather agrees on success across ranks internally and raises on ALL ranks
            if mode == "overlap":
                agather = AsyncRootGather(T, E, Np, info_w, dev)
            elif WL["family"] == "ww":
                peer = PeerGather(eng, T, Np, mode="all" if mode == "fused-all" else "root")
            else:
                raise RuntimeError("fused exchange is implemented for Waterworld only")
        except Exception as ex:   # e.g. no peer access between the GPUs of this box
            peer = agather = None
            mode, note = "nccl", "%s unavailable (%s)" % (mode, type(ex).__name__)
    packed = PackedTrajectory(T, E, Np, info_w, dev)
    obs_buf = torch.empty((T, E, Np, D), device=dev)
    out = (obs_buf, packed.rew, packed.done, packed.info)
    g_obs = None
    if world > 1 and a.gather_obs:
        g_obs = torch.empty((world * T, E, Np, D), device=dev)
    step_counter = [0]

    def run_rollout(i):
        k = step_counter[0]
        step_counter[0] += 1
        if agather is not None:
            agather.before_reuse(k)
            rew_b, done_b, info_b = agather.buffers(k)
            eng.rollout(actions[i % n_act], auto_reset=True, out=(obs_buf, rew_b, done_b, info_b))
        else:
            if peer is not None:
                peer.arm(k)
            eng.rollout(actions[i % n_act], auto_reset=True, out=out)
        return k

    def exchange(k):
        if world > 1:
            if agather is not None:
                agather.submit(k)
            elif peer is not None:
                peer.complete()
            else:
                packed.gather_raw()
            if a.gather_obs:
                dist.all_gather_into_tensor(g_obs, obs_buf)

    def one_step(i):
        exchange(run_rollout(i))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(3, a.warmup)):
        one_step(i)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    ev0.record()
    for i in range(a.steps):
        kev[i][0].record()
        k_ = run_rollout(i)
        kev[i][1].record()
        exchange(k_)
    if agather is not None:   # the timed region ends when the last exchange has landed
        torch.cuda.current_stream(dev).wait_stream(agather.comm)
    ev1.record()
    barrier()
    launches = launch_count() - l0
    ms = ev0.elapsed_time(ev1)
    # The timed region is only a few ms long, shorter than one nvidia-smi poll: keep the identical
    # workload running for ~0.8 s more (untimed) so the clock/throttle record is taken under load.
    if sampler:
        sampler.samples.clear()
    t_load = time.perf_counter()
    i = 0
    while time.perf_counter() - t_load < 0.8:
        for _ in range(20):
            eng.rollout(actions[i % n_act], auto_reset=True, out=out)
            i += 1
        torch.cuda.synchronize()
    for x in (peer, agather):
        if x is not None:
            x.close()
    kern_ms = float(np.mean([s.elapsed_time(e) for s, e in kev]))
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    agent_steps = world * E * Np * T * a.steps
    value = agent_steps / (ms * 1e-3)

    # ---- e2e: the same rollout through the host-buffer C-ABI entry point (pinned host tensors)
    e2e = None
    if not a.no_e2e:
        Te = min(T, 16)
        Te = min(Te, max(1, (1 << 28) // (E * Np * D * 4)))      # <= 256 MB of obs per call
        h_act = [make_actions(a.workload, Te, E, dev, g, host=True).pin_memory() for _ in range(2)]
        h_out = (torch.empty((Te, E, Np, D)).pin_memory(), torch.empty((Te, E, Np)).pin_memory(),
                 torch.empty((Te, E), dtype=torch.uint8).pin_memory(),
                 torch.empty((Te, E, 2) if info_w == 2 else (Te, E), dtype=torch.int32).pin_memory())
        for i in range(2):
            eng.rollout_host(h_act[i % 2], *h_out)
        barrier()
        n_e2e = max(3, min(a.steps, 10))
        t0 = time.perf_counter()
        for i in range(n_e2e):
            eng.rollout_host(h_act[i % 2], *h_out)     # returns after results are in host memory
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * E * Np * Te * n_e2e / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(h_act[0].numel() * h_act[0].element_size()),
               "d2h_bytes_per_step": int(sum(x.numel() * x.element_size() for x in h_out)),
               "t_inner": Te, "api": "madrl_%s_rollout_host (pinned host buffers)" % {"ww": "ww", "pe": "pursuit", "hw": "hostage"}[WL["family"]]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    bpe = bytes_per_env_step(a.workload)
    peak, peak_kind = measured_peak_gbs()
    achieved = bpe * E * T / (kern_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": max(3, a.warmup), "ms_per_step": ms / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": ("u8/int32 + f32 obs" if WL["family"] == "pe" else "f32"),
        "data": "synthetic",
        "config": {"workload": WL["desc"] + ", %d envs per GPU, %d lockstep env steps per rollout launch, "
                               "auto-reset (VecEnvExecutor semantics)" % (E, T),
                   "envs_per_gpu": E, "t_inner": T,
                   "actions": ("uniform {0..4}" if WL["family"] == "pe" else "0.5*N(0,1)") + ", HBM-resident",
                   "l2": "outputs per launch (%.0f MB) exceed L2" % (out[0].numel() * 4 / 1e6),
                   "parallelism": "env-shard x%d" % world,
                   "gather": {"none": "none",
                              "overlap": "packed rew/done/info copied to rank 0 over NVLink by the copy engines on "
                                         "a side stream + 1 completion all-reduce per rollout, overlapped with the "
                                         "next rollout",
                              "nccl": "one packed NCCL all_gather of rew/done/info per rollout",
                              "fused": "rew/done/info rows stored by the rollout kernel into rank 0's gather "
                                       "buffers over NVLink peer memory + 1 completion all-reduce per rollout",
                              "fused-all": "rew/done/info rows stored by the rollout kernel into every rank's "
                                           "gather buffers over NVLink peer memory + 1 completion all-reduce"}[mode]
                             + (" [%s]" % note if note else "") + (" + obs all_gather" if a.gather_obs else "")},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the ncu --set full
                     # captures summarised in profiles/r1_{ww,pe,hw}_kernel_final_full.md, expressed per
                     # env-step (Waterworld C2: 4.537 GB per 4096x256 launch; Pursuit C3: 2.581 GB per
                     # 65536x8; Hostage C5: 1.637 GB per 8192x32) and scaled to this launch
                     "traffic": ({"waterworld": 4327.0, "pursuit": 4923.0, "hostage": 6246.0}[a.workload] * E * T
                                 if a.workload in ("waterworld", "pursuit", "hostage") else None),
                     "traffic_unit": "bytes per launch (algorithmic: %d)" % (bpe * E * T),
                     "peak_kind": peak_kind,
                     "kernel": {"ww": "ww_kernel<float>", "pe": "pe_kernel", "hw": "hw_kernel<float>"}[WL["family"]],
                     "kernel_ms": kern_ms,
                     "algorithmic_bytes_per_env_step": bpe},
        "gpu_launches": int(launches),
        "clocks": dict(sampler.summary(), window="same rollout workload kept running for 0.8 s right "
                       "after the timed region (the timed region itself is shorter than one poll)")
        if sampler else None,
    }
    if e2e:
        line["e2e"] = e2e
    if world == 1 and not a.no_cpu:
        procs = host_cores
        v = cpu_baseline(a.cpu_seconds, procs, a.workload)
        v1 = cpu_baseline(min(4.0, a.cpu_seconds), 1, a.workload)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": procs, "kind": "port",
                                "single_core_value": v1,
                                "sample": "%d processes x %.0f s of step() (one env each), float64 NumPy "
                                          "oracle port of the reference step()" % (procs, a.cpu_seconds)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
