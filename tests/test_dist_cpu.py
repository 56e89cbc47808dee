"""Multi-process (world_size 2, gloo, CPU) test of the sharding / gather host logic: each rank
steps ITS shard of the env batch with the oracle (keyed by global env id, exactly like the CUDA
engines), the trajectory tensors are gathered once per rollout, and the result must equal the
unsharded batch."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from madrl_b200.dist import gather_trajectories, shard_range

E_GLOBAL, T, NP, SEED = 6, 5, 3, 42
CFG = dict(n_pursuers=NP, n_evaders=3, n_poison=2, n_sensors=6, n_coop=1, radius=0.05)


def rollout_oracle(lo, hi):
    from oracle.philox import Stream
    from oracle.waterworld_oracle import WaterworldOracle
    acts = np.random.RandomState(0).randn(T, E_GLOBAL, NP, 2) * 0.5
    obs = np.zeros((T, hi - lo, NP, 6 * 7 + 3), dtype=np.float32)
    rew = np.zeros((T, hi - lo, NP), dtype=np.float32)
    for e in range(lo, hi):
        o = WaterworldOracle(rng=Stream(SEED, e), **CFG)   # key = (seed, GLOBAL env id)
        o.reset()
        for t in range(T):
            ob, r, _, _ = o.step(acts[t, e])
            obs[t, e - lo] = np.array(ob)
            rew[t, e - lo] = r
    return torch.from_numpy(obs), torch.from_numpy(rew)


def _worker(rank, world, port, ragged, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if ragged:
            lo, hi = (0, 4) if rank == 0 else (4, E_GLOBAL)
        else:
            lo, hi = shard_range(E_GLOBAL, rank, world)
        obs, rew = rollout_oracle(lo, hi)
        g_obs, g_rew = gather_trajectories((obs, rew), env_dim=1, equal_shards=not ragged)
        d_obs, = gather_trajectories((obs,), env_dim=1, dst=0) if not ragged else (g_obs,)
        if not ragged:     # the single-collective packed exchange used by bench.py
            from madrl_b200.dist import PackedTrajectory
            pk = PackedTrajectory(T, hi - lo, NP, 2, "cpu")
            pk.rew.copy_(rew)
            pk.done.fill_(rank)
            pk.info.fill_(7 + rank)
            p_rew, p_done, p_info = pk.gather()
            assert torch.equal(torch.cat(list(p_rew), dim=1), g_rew)
            assert [int(p_done[w].max()) for w in range(world)] == list(range(world))
            assert [int(p_info[w].min()) for w in range(world)] == [7 + w for w in range(world)]
        if rank == 0:
            f_obs, f_rew = rollout_oracle(0, E_GLOBAL)
            ok = torch.equal(g_obs, f_obs) and torch.equal(g_rew, f_rew) and torch.equal(d_obs, f_obs)
            q.put(bool(ok))
        elif not ragged:
            assert d_obs is None
    finally:
        dist.destroy_process_group()


def _ragged_peer_worker(rank, world, port, q):
    """the peer-memory exchanges index the root's buffers by rank * local size, so they
    must refuse ragged shards on EVERY rank (here: the rank-consistency check they call first)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from madrl_b200.dist import _same_on_all_ranks
        _same_on_all_ranks((5, 3), "shape")                       # equal: passes
        try:
            _same_on_all_ranks((2 + rank, 3), "(n_envs, n_agents)")   # 10 envs on 4 ranks -> 2,3,2,3
            q.put(False)
        except ValueError as ex:
            q.put("differs across ranks" in str(ex))
    finally:
        dist.destroy_process_group()


def test_peer_exchange_refuses_ragged_shards():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_ragged_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() is True and q.get() is True


def test_packed_trajectory_sections_are_16_byte_aligned():
    """the kernels store info rows as 8-byte words; with T*E*A odd (fp32) the old layout put
    the info section at 4 mod 8."""
    from madrl_b200.dist import PackedTrajectory
    for T, E, A, w in ((3, 1, 5, 2), (1, 1, 1, 2), (7, 3, 3, 1), (5, 2, 3, 2)):
        pk = PackedTrajectory(T, E, A, w, "cpu", obs_dim=11, act_shape=(A, 2))
        for k, (off, n) in pk.offsets.items():
            assert off % 16 == 0, (k, off)
            assert getattr(pk, k).data_ptr() - pk.buf.data_ptr() == off
        assert pk.rew.shape == (T, E, A) and pk.obs.shape == (T, E, A, 11) and pk.act.shape == (T, E, A, 2)
        assert pk.info.shape == ((T, E, w) if w > 1 else (T, E)) and pk.nbytes % 16 == 0
        pk.buf.zero_()
        pk.info.fill_(-1)
        pk.done.fill_(3)
        assert float(pk.rew.abs().sum()) == 0 and float(pk.obs.abs().sum()) == 0   # no overlap


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(ragged):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ragged, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() is True


def test_shard_range_partitions():
    for E in (1, 7, 4096, 32768):
        for W in (1, 2, 3, 8):
            rs = [shard_range(E, r, W) for r in range(W)]
            assert rs[0][0] == 0 and rs[-1][1] == E
            assert all(rs[i][1] == rs[i + 1][0] for i in range(W - 1))
            assert max(h - l for l, h in rs) - min(h - l for l, h in rs) <= 1


def test_two_rank_gloo_gather_equals_unsharded():
    _run(ragged=False)


def test_two_rank_gloo_ragged_gather():
    _run(ragged=True)
