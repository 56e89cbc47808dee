"""Env-axis sharding across the GPUs of one box.

Env instances are independent, so the batch shards trivially: rank r owns the contiguous global
env range ``shard_range(E, r, W)`` and builds its engine with ``env_id_base = lo`` -- the RNG key
is (seed, GLOBAL env id), so a sharded batch reproduces the unsharded one bit for bit.  There is
no collective inside the step loop; once per rollout the trajectory tensors are exchanged with
ONE ``all_gather`` (or gather-to-rank) per tensor over NCCL/NVLink (``gloo`` on CPU in the tests).
This is the B200 counterpart of the reference's pickled-path return from sampler workers
(rllab/rllab/sampler/stateful_pool.py:102-157, rltools/rltools/samplers/parallel.py:214-222).
"""
import os

import torch
import torch.distributed as dist


def world_info():
    """(rank, world_size, local_rank) from torchrun's environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(n_envs_global, rank, world):
    """Contiguous, balanced partition of [0, n_envs_global): returns (lo, hi) for `rank`."""
    lo = (n_envs_global * rank) // world
    hi = (n_envs_global * (rank + 1)) // world
    return lo, hi


def make_sharded(engine_cls, n_envs_global, *args, rank=None, world=None, **kwargs):
    """Build this rank's shard of a Batched* engine (env ids [lo, hi) of the global batch)."""
    r, w, _ = world_info()
    rank = r if rank is None else rank
    world = w if world is None else world
    lo, hi = shard_range(n_envs_global, rank, world)
    if hi <= lo:
        raise ValueError("rank %d of %d owns no envs of a %d-env batch" % (rank, world, n_envs_global))
    eng = engine_cls(hi - lo, *args, env_id_base=kwargs.pop("env_id_base", 0) + lo, **kwargs)
    eng.shard = (lo, hi, n_envs_global)
    return eng


def gather_trajectories(tensors, env_dim=1, group=None, dst=None, equal_shards=True):
    """All-gather (or gather to `dst`) per-rank trajectory tensors along the env axis.

    `tensors`: tuple of tensors shaped [T, E_local, ...] (env_dim=1) or [E_local, ...] (env_dim=0).
    Returns a tuple of tensors shaped [..., E_global, ...] in global env order (None on ranks other
    than `dst` when gathering to one rank).  One collective per tensor, no per-step traffic.
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tuple(tensors)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    out = []
    for t in tensors:
        t = t.contiguous()
        if equal_shards:
            if dst is None:
                # concatenated-along-dim-0 output form (accepted by both NCCL and gloo)
                buf = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
                dist.all_gather_into_tensor(buf, t, group=group)
                parts = list(buf.view((world,) + tuple(t.shape)).unbind(0))
            else:
                parts = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
                dist.gather(t, parts, dst=dst, group=group)
        else:  # ragged shards: exchange sizes, pad to the largest shard, gather, trim
            n = torch.tensor([t.shape[env_dim]], dtype=torch.int64, device=t.device)
            sizes = torch.empty(world, dtype=torch.int64, device=t.device)
            dist.all_gather_into_tensor(sizes, n, group=group)
            sizes = [int(v) for v in sizes.tolist()]
            shp = list(t.shape)
            shp[env_dim] = max(sizes)
            padded = torch.zeros(shp, dtype=t.dtype, device=t.device)
            padded.narrow(env_dim, 0, t.shape[env_dim]).copy_(t)
            buf = torch.empty((world * shp[0],) + tuple(shp[1:]), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(buf, padded, group=group)
            parts = [c.narrow(env_dim, 0, sz) for c, sz in zip(buf.view([world] + shp).unbind(0), sizes)]
        out.append(torch.cat(parts, dim=env_dim) if parts is not None else None)
    return tuple(out)


class PackedTrajectory(object):
    """Reward / done / info tensors of one rollout carved out of ONE contiguous byte buffer, so the
    per-rollout exchange is a single collective (`gather()`), not one per tensor.  The env kernels
    write straight into the views."""

    def __init__(self, T, n_envs, n_agents, info_width, device, rew_dtype=torch.float32):
        esz = torch.empty((), dtype=rew_dtype).element_size()
        self.shapes = dict(rew=(T, n_envs, n_agents), info=(T, n_envs, info_width) if info_width > 1 else (T, n_envs),
                           done=(T, n_envs))
        r_b = T * n_envs * n_agents * esz
        i_b = T * n_envs * info_width * 4
        d_b = T * n_envs
        self.nbytes = (r_b + i_b + d_b + 15) // 16 * 16
        self.buf = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        self.rew = self.buf[:r_b].view(rew_dtype).view(self.shapes['rew'])
        self.info = self.buf[r_b:r_b + i_b].view(torch.int32).view(self.shapes['info'])
        self.done = self.buf[r_b + i_b:r_b + i_b + d_b].view(self.shapes['done'])
        self._r_b, self._i_b, self._d_b, self._rew_dtype = r_b, i_b, d_b, rew_dtype
        self._gbuf = None

    def gather(self, group=None):
        """ONE all_gather of the packed buffer; returns (rew, done, info) with a leading rank axis
        ([W, T, E_local, ...], unpacked copies of the per-rank slices)."""
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return self.rew.unsqueeze(0), self.done.unsqueeze(0), self.info.unsqueeze(0)
        world = dist.get_world_size(group)
        if self._gbuf is None:
            self._gbuf = torch.empty(world * self.nbytes, dtype=torch.uint8, device=self.buf.device)
        dist.all_gather_into_tensor(self._gbuf, self.buf, group=group)
        g = self._gbuf.view(world, self.nbytes)
        r_b, i_b, d_b = self._r_b, self._i_b, self._d_b
        rew = torch.stack([g[w, :r_b].view(self._rew_dtype).view(self.shapes['rew']) for w in range(world)])
        info = torch.stack([g[w, r_b:r_b + i_b].view(torch.int32).view(self.shapes['info']) for w in range(world)])
        done = torch.stack([g[w, r_b + i_b:r_b + i_b + d_b].view(self.shapes['done']) for w in range(world)])
        return rew, done, info

    def gather_raw(self, group=None):
        """The collective only (no unpacking): what sits in the benchmark's timed region."""
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            world = dist.get_world_size(group)
            if self._gbuf is None:
                self._gbuf = torch.empty(world * self.nbytes, dtype=torch.uint8, device=self.buf.device)
            dist.all_gather_into_tensor(self._gbuf, self.buf, group=group)
        return self._gbuf


class _DevBuf(object):
    """A raw device allocation exposed to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr, nbytes):
        self.ptr, self.nbytes = ptr, nbytes
        self.__cuda_array_interface__ = dict(shape=(nbytes,), typestr="|u1", data=(ptr, False), version=2)


class PeerGather(object):
    """Fused per-rollout exchange over NVLink peer memory (no separate collective pass).

    The receiving ranks (``mode='root'``: rank ``root`` only, like the reference's master process
    that collects the workers' paths; ``mode='all'``: every rank) allocate double-buffered gather
    buffers with ``madrl_ipc_alloc`` in an env-major layout

        rew [W, E, Tmax, A]     done [W, E, Tmax]     info [W, E, Tmax, 2]

    The 64-byte CUDA IPC handles are exchanged with ``all_gather_object``; every rank opens the
    destination buffers with its own device current (``madrl_ipc_open``) and hands the mappings to
    the engine (``madrl_ww_set_peers``).  The rollout kernel then writes each env's reward / done /
    info rows into slot ``rank`` of every destination buffer while it computes (rows staged in
    registers, coalesced runs over NVLink).  What is left of the "gather" is ``complete()``: one tiny
    all-reduce that orders every rank's kernel completion before the buffers are read.  The two
    buffer sets alternate so that a fast rank writing rollout k+1 never touches the set a slow rank
    is still reading for rollout k.
    """

    def __init__(self, engine, t_max, n_agents, info_width=2, n_sets=2, group=None, mode="root", root=0):
        import ctypes as C
        from . import _lib
        assert dist.is_initialized() and mode in ("root", "all")
        self._L = _lib.lib()
        self.engine, self.group, self.t_max = engine, group, t_max
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.receivers = [root] if mode == "root" else list(range(self.world))
        self.is_receiver = self.rank in self.receivers
        dev, E, W = engine.device, engine.n_envs, self.world
        esz = torch.empty((), dtype=engine.dtype).element_size()
        r_b = W * t_max * E * n_agents * esz
        d_b = (W * t_max * E + 255) // 256 * 256
        i_b = W * t_max * E * info_width * 4
        nbytes = r_b + i_b + d_b

        def views(buf_u8):
            return (buf_u8[:r_b].view(engine.dtype).view(W, E, t_max, n_agents),
                    buf_u8[r_b + i_b:r_b + i_b + W * t_max * E].view(W, E, t_max),
                    buf_u8[r_b:r_b + i_b].view(torch.int32).view(W, E, t_max, info_width))

        self.sets, self._dest, self._own, self._opened = [], [], [], []
        with torch.cuda.device(dev):
            for _ in range(n_sets):
                ptr, handle = C.c_void_p(), C.create_string_buffer(64)
                if self.is_receiver:
                    _lib.check(self._L.madrl_ipc_alloc(nbytes, C.byref(ptr), handle))
                    self._own.append(ptr.value)
                everyone = [None] * W
                dist.all_gather_object(everyone, handle.raw if self.is_receiver else None, group=group)
                dest = ([], [], [])
                local = None
                for r in self.receivers:
                    if r == self.rank:
                        base = ptr.value
                    else:
                        q = C.c_void_p()
                        _lib.check(self._L.madrl_ipc_open(everyone[r], C.byref(q)))
                        self._opened.append(q.value)
                        base = q.value
                    v = views(torch.as_tensor(_DevBuf(base, nbytes), device=dev))
                    if r == self.rank:
                        local = v
                    for k in range(3):
                        dest[k].append(v[k])
                self.sets.append(local)
                self._dest.append(dest)
            torch.cuda.synchronize(dev)
        dist.barrier(group=group)
        self._flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self._cur = -1

    def arm(self, k):
        """Direct the next rollout's exchange at buffer set k % n_sets."""
        self._cur = k % len(self.sets)
        pr, pd, pi = self._dest[self._cur]
        self.engine.set_peers(self.rank, self.t_max, pr, pd, pi)

    def complete(self):
        """Order all ranks' rollout kernels before the gathered buffers are read (stream-ordered).
        On a receiving rank returns (rew [W,E,Tmax,A], done [W,E,Tmax], info [W,E,Tmax,2]) of the
        armed set (slot r = rank r's envs); None elsewhere."""
        dist.all_reduce(self._flag, group=self.group)
        return self.sets[self._cur]

    def close(self):
        self.engine.clear_peers()
        torch.cuda.synchronize(self.engine.device)
        dist.barrier(group=self.group)
        with torch.cuda.device(self.engine.device):
            for q in self._opened:
                self._L.madrl_ipc_close(q)
            dist.barrier(group=self.group)
            for q in self._own:
                self._L.madrl_ipc_free(q)
        self._opened, self._own, self._dest, self.sets = [], [], [], []


class AsyncRootGather(object):
    """Per-rollout gather of the packed reward / done / info buffers to a root rank that OVERLAPS
    with the next rollout: after rollout k every rank enqueues, on a side stream, ONE device-to-device
    copy of its packed buffer into its slot of the root's (CUDA-IPC mapped) gather buffer -- a copy
    engine transfer over NVLink that uses no SMs, so the persistent rollout kernel of step k+1 runs
    undisturbed -- followed by one tiny completion all-reduce.  This is the reference's
    "workers return their paths to the master" (rllab/rllab/sampler/stateful_pool.py:102-157) as a
    B200 NVLink pattern.  Local buffers and root buffers are double-buffered.
    """

    def __init__(self, T, n_envs, n_agents, info_width, device, rew_dtype=torch.float32, root=0,
                 n_sets=2, group=None):
        import ctypes as C
        from . import _lib
        assert dist.is_initialized()
        self._L = _lib.lib()
        self.group, self.root, self.device = group, root, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.local = [PackedTrajectory(T, n_envs, n_agents, info_width, device, rew_dtype) for _ in range(n_sets)]
        nb = self.local[0].nbytes
        self._slots, self._own, self._opened, self.root_bufs = [], [], [], []
        failure = None
        with torch.cuda.device(device):
            for _ in range(n_sets):
                ptr, handle = C.c_void_p(), C.create_string_buffer(64)
                try:
                    if self.rank == root:
                        _lib.check(self._L.madrl_ipc_alloc(self.world * nb, C.byref(ptr), handle))
                        self._own.append(ptr.value)
                except Exception as ex:          # keep going: every rank must reach the collectives below
                    failure = failure or ex
                everyone = [None] * self.world
                dist.all_gather_object(everyone, handle.raw if (self.rank == root and failure is None) else None,
                                       group=group)
                try:
                    if failure is None and everyone[root] is None:
                        raise RuntimeError("root could not export its gather buffer")
                    if failure is None:
                        if self.rank == root:
                            base = ptr.value
                        else:
                            q = C.c_void_p()
                            _lib.check(self._L.madrl_ipc_open(everyone[root], C.byref(q)))
                            self._opened.append(q.value)
                            base = q.value
                        whole = torch.as_tensor(_DevBuf(base, self.world * nb), device=device)
                        self.root_bufs.append(whole if self.rank == root else None)
                        self._slots.append(whole[self.rank * nb:(self.rank + 1) * nb])
                except Exception as ex:
                    failure = failure or ex
            torch.cuda.synchronize(device)
        # agree on success: either every rank uses the peer-memory path or none does
        ok = torch.tensor([0.0 if failure is not None else 1.0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if ok.item() == 0:
            self._release()
            raise RuntimeError("AsyncRootGather unavailable: %r" % (failure or "failure on another rank"))
        self.comm = torch.cuda.Stream(device=device)
        self._flag = torch.zeros(1, dtype=torch.int32, device=device)
        self._ev = [torch.cuda.Event() for _ in range(n_sets)]

    def buffers(self, k):
        """(rew, done, info) views the rollout k must write into."""
        p = self.local[k % len(self.local)]
        return p.rew, p.done, p.info

    def submit(self, k):
        """Enqueue the exchange of rollout k (already launched on the current stream)."""
        i = k % len(self.local)
        self._ev[i].record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self._ev[i])
            self._slots[i].copy_(self.local[i].buf, non_blocking=True)     # copy engine, NVLink P2P
            dist.all_reduce(self._flag, group=self.group)                 # completion ordering
            self._ev[i].record(self.comm)

    def before_reuse(self, k):
        """Make the current stream wait until the exchange that last used buffer set k%n is done."""
        torch.cuda.current_stream(self.device).wait_event(self._ev[k % len(self.local)])

    def result(self, k):
        """On the root: [W, nbytes] uint8 view of the gathered packed buffers of rollout k (after
        `before_reuse(k)` / a synchronize); unpack with PackedTrajectory offsets.  None elsewhere."""
        if self.rank != self.root:
            return None
        return self.root_bufs[k % len(self.local)].view(self.world, self.local[0].nbytes)

    def _release(self):
        with torch.cuda.device(self.device):
            for q in self._opened:
                self._L.madrl_ipc_close(q)
            dist.barrier(group=self.group)
            for q in self._own:
                self._L.madrl_ipc_free(q)
        self._opened, self._own = [], []

    def close(self):
        self.comm.synchronize()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        self._release()
