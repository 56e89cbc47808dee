"""Teacher-forced single-step comparison of the fp32 kernels with the float64 oracle / the real
reference's recorded states (TEST INFRASTRUCTURE shared by the GPU tests, the emulated CPU tests and
``__graft_entry__.smoke()``).

Every transition is classified with ``oracle/fragility.py``: a transition whose DYNAMICS contain a
comparison within EPS of its threshold is excluded (and counted); otherwise events, rewards and the
post-state must agree, and every observation element whose own sensing predicates are not within
EPS must agree within TOL (1e-5, the tolerance BASELINE.json's north_star states).  What was
checked / excluded is returned in an ``ExclusionLog``.

Engines are driven through small adapters (`read()` -> dict of NumPy arrays named like
``Batched*.state``; `write(dict)`; `step(actions)`), one for the torch/CUDA engines and one for the
emulator engines of tests/emu.
"""
import numpy as np

from oracle.fragility import ExclusionLog, hw_fragility, ww_fragility
from oracle.hostage_oracle import HostageOracle
from oracle.philox import Stream
from oracle.waterworld_oracle import WaterworldOracle

TOL32 = 1e-5
EPS_FRAGILE = 3e-7


# ------------------------------------------------------------------------------------ adapters
class TorchAdapter(object):
    """madrl_b200.Batched{MAWaterWorld,HostageWorld} (CUDA)."""

    def __init__(self, eng):
        self.eng, self.E = eng, eng.n_envs

    def read(self):
        return {k: v.cpu().numpy().copy() for k, v in self.eng.state.items()}

    def write(self, d):
        import torch
        st = self.eng.state
        for k, v in d.items():
            st[k].copy_(torch.as_tensor(np.ascontiguousarray(v)).to(st[k].dtype))

    def step(self, act):
        import torch
        a = torch.as_tensor(act[None])
        obs, rew, done, info = [x.cpu().numpy() for x in self.eng.rollout(a, auto_reset=False)]
        return obs[0], rew[0], done[0], info[0]


class EmuAdapter(object):
    """tests/emu/driver.py Emu{Waterworld,Hostage} (kernel source run on the CPU)."""

    def __init__(self, eng, family):
        self.eng, self.E, self.family = eng, eng.n_envs, family

    def _views(self):
        g, Ly, E = self.eng, self.eng.layout, self.eng.n_envs
        N = int(Ly.n_obj)
        objs = g.view(Ly.objs, g.obs_dtype, (E, 4, N))
        d = dict(pos_x=objs[:, 0], pos_y=objs[:, 1], vel_x=objs[:, 2], vel_y=objs[:, 3],
                 timestep=g.view(Ly.timestep, np.int32, (E,)),
                 rng_counter=g.view(Ly.rng_counter, np.int64, (E,)))
        if self.family == "ww":
            obst = g.view(Ly.obst, g.obs_dtype, (E, 2))
            d.update(obst_x=obst[:, 0], obst_y=obst[:, 1])
        else:
            fixed = g.view(Ly.fixed, g.obs_dtype, (E, 4))
            d.update(key=fixed[:, 0:2], bomb=fixed[:, 2:4], flags=g.view(Ly.flags, np.int32, (E,)),
                     saved=g.view(Ly.saved, np.uint8, (E, g.n_hostages)))
        return d

    def read(self):
        return {k: np.array(v) for k, v in self._views().items()}

    def write(self, d):
        v = self._views()
        for k, x in d.items():
            v[k][...] = x

    def step(self, act):
        obs, rew, done, info = self.eng.rollout(act[None], auto_reset=False)
        return obs[0], rew[0], done[0], info[0]


# ------------------------------------------------------------------------------------ Waterworld
def ww_state_of(arr, e, Np, Ne):
    X = np.stack([arr['pos_x'][e], arr['pos_y'][e]], 1).astype(np.float64)
    V = np.stack([arr['vel_x'][e], arr['vel_y'][e]], 1).astype(np.float64)
    return dict(px=X[:Np], pv=V[:Np], ex=X[Np:Np + Ne], ev=V[Np:Np + Ne], ox=X[Np + Ne:], ov=V[Np + Ne:],
                obst=np.array([[arr['obst_x'][e], arr['obst_y'][e]]], dtype=np.float64),
                t=int(arr['timestep'][e]), counter=int(arr['rng_counter'][e]))


def _ww_compare(orc, pre, act, got, post, log, where):
    """One transition: oracle from `pre` with `act` vs the engine's (obs, rew, done, info) + post."""
    obs, rew, done, info = got
    log.add('transitions')
    dyn, ok = ww_fragility(orc, pre, act, EPS_FRAGILE)
    if dyn:
        log.add('excluded_dyn_fragile')
        return False
    orc.set_state(pre)
    oo, rr, dd, ii = orc.step(np.asarray(act, np.float64))
    oo = np.array(oo)
    assert [ii['evcatches'], ii['pocatches']] == [int(info[0]), int(info[1])], where
    assert bool(done) == dd, where
    err = np.abs(oo - obs)
    log.add('obs_elements', ok.size)
    log.add('obs_elements_compared', int(ok.sum()))
    log.err('max_abs_err_obs', err[ok].max())
    assert err[ok].max() <= TOL32, (where, float(err[ok].max()), np.argwhere(ok & (err > TOL32))[:4])
    assert np.abs(rr - rew).max() <= TOL32, where
    assert post['counter'] == orc.np_random.counter, where
    for k in ('px', 'pv', 'ex', 'ev', 'ox', 'ov'):
        d = np.abs(post[k] - getattr(orc, k)).max()
        log.err('max_abs_err_state', d)
        assert d <= TOL32, (where, k, d)
    log.add('checked')
    return True


def ww_self_teacher_forced(ad, cfg, seed, T, std, case, env_ids=None, env_id_base=0):
    """Single-step teacher forcing from the engine's OWN fp32 states (all envs, or `env_ids`)."""
    Np, Ne = cfg['n_pursuers'], cfg['n_evaders']
    rs = np.random.RandomState(3)
    orc = WaterworldOracle(rng=Stream(seed, 0), **cfg)
    log = ExclusionLog(case, family="waterworld", source="engine fp32 states", E=ad.E, T=T, eps=EPS_FRAGILE, tol=TOL32)
    ids = range(ad.E) if env_ids is None else env_ids
    for t in range(T):
        act = (rs.randn(ad.E, Np, 2) * std).astype(np.float32)
        arr = ad.read()
        obs, rew, done, info = ad.step(act)
        arr2 = ad.read()
        for e in ids:
            pre = ww_state_of(arr, e, Np, Ne)
            orc.np_random = Stream(seed, env_id_base + e, counter=pre['counter'])
            _ww_compare(orc, pre, act[e], (obs[e], rew[e], done[e], info[e]), ww_state_of(arr2, e, Np, Ne),
                        log, (case, t, e))
    return log


def ww_golden_teacher_forced(ad, g, cfg, case):
    """Teacher forcing from the REAL reference's recorded float64 states (cast to the engine's
    dtype), compared with the reference's recorded outputs of the same step."""
    Np, Ne = cfg['n_pursuers'], cfg['n_evaders']
    seed, env_id = int(g['seed']), int(g['env_id'])
    T = g['actions'].shape[0]
    orc = WaterworldOracle(rng=Stream(seed, env_id), **cfg)
    log = ExclusionLog(case, family="waterworld", source="reference float64 states (tests/golden)", E=1, T=T,
                       eps=EPS_FRAGILE, tol=TOL32)
    for t in range(T):
        pos, vel = g['st_pos'][t], g['st_vel'][t]
        ad.write(dict(pos_x=pos[None, :, 0], pos_y=pos[None, :, 1], vel_x=vel[None, :, 0], vel_y=vel[None, :, 1],
                      obst_x=g['st_obst'][t][None, 0], obst_y=g['st_obst'][t][None, 1],
                      timestep=np.array([g['st_t'][t]], np.int32),
                      rng_counter=np.array([g['st_counter'][t]], np.int64)))
        act = g['actions'][t].astype(np.float32)
        obs, rew, done, info = ad.step(act[None])
        log.add('transitions')
        ref = dict(px=pos[:Np], pv=vel[:Np], ex=pos[Np:Np + Ne], ev=vel[Np:Np + Ne], ox=pos[Np + Ne:],
                   ov=vel[Np + Ne:], obst=g['st_obst'][t][None], t=int(g['st_t'][t]), counter=int(g['st_counter'][t]))
        dyn, ok = ww_fragility(orc, ref, act, EPS_FRAGILE)
        if dyn:
            log.add('excluded_dyn_fragile')
            continue
        assert list(info[0]) == list(g['info'][t]), (case, t)
        assert bool(done[0]) == bool(g['done'][t]), (case, t)
        err = np.abs(g['obs'][t] - obs[0])
        log.add('obs_elements', ok.size)
        log.add('obs_elements_compared', int(ok.sum()))
        log.err('max_abs_err_obs', err[ok].max())
        assert err[ok].max() <= TOL32, (case, t, float(err[ok].max()))
        assert np.abs(g['rew'][t] - rew[0]).max() <= TOL32, (case, t)
        post = ad.read()
        assert int(post['rng_counter'][0]) == int(g['st_counter'][t + 1]), (case, t)
        for got, want in ((post['pos_x'][0], g['st_pos'][t + 1][:, 0]), (post['pos_y'][0], g['st_pos'][t + 1][:, 1]),
                          (post['vel_x'][0], g['st_vel'][t + 1][:, 0]), (post['vel_y'][0], g['st_vel'][t + 1][:, 1])):
            d = np.abs(got.astype(np.float64) - want).max()
            log.err('max_abs_err_state', d)
            assert d <= TOL32, (case, t, d)
        log.add('checked')
    return log


# ------------------------------------------------------------------------------------ Hostage
def hw_state_of(arr, e, Nr, Nc):
    X = np.stack([arr['pos_x'][e], arr['pos_y'][e]], 1).astype(np.float64)
    V = np.stack([arr['vel_x'][e], arr['vel_y'][e]], 1).astype(np.float64)
    f = int(arr['flags'][e])
    return dict(rx=X[:Nr], rv=V[:Nr], cx=X[Nr:Nr + Nc], cv=V[Nr:Nr + Nc], hx=X[Nr + Nc:],
                key=arr['key'][e].astype(np.float64)[None], bomb=arr['bomb'][e].astype(np.float64)[None],
                saved=arr['saved'][e].astype(bool), gate_open=bool(f & 1), bombed=bool(f & 2),
                t=int(arr['timestep'][e]), counter=int(arr['rng_counter'][e]))


def hw_self_teacher_forced(ad, args, kw, seed, T, std, case, reset_done, env_ids=None):
    """`reset_done(mask)` re-initialises finished envs so that live ones keep being stepped."""
    Nr, Nh, Nc = args[0], args[1], args[2]
    rs = np.random.RandomState(3)
    orc = HostageOracle(*args, rng=Stream(seed, 0), **kw)
    log = ExclusionLog(case, family="hostage", source="engine fp32 states", E=ad.E, T=T, eps=EPS_FRAGILE, tol=TOL32,
                       excluded_terminal_pre_state=0)
    ids = range(ad.E) if env_ids is None else env_ids
    for t in range(T):
        act = (rs.randn(ad.E, Nr, 2) * std).astype(np.float32)
        arr = ad.read()
        obs, rew, done, info = ad.step(act)
        arr2 = ad.read()
        for e in ids:
            pre = hw_state_of(arr, e, Nr, Nc)
            log.add('transitions')
            if pre['bombed'] or pre['saved'].all():     # stepping a finished env is outside the contract
                log.add('excluded_terminal_pre_state')
                log.add('excluded_other')
                continue
            dyn, ok = hw_fragility(orc, pre, act[e], EPS_FRAGILE)
            if dyn:
                log.add('excluded_dyn_fragile')
                continue
            orc.np_random = Stream(seed, e, counter=pre['counter'])
            orc.set_state(pre)
            oo, rr, dd, ii = orc.step(act[e].astype(np.float64))
            oo = np.array(oo)
            assert [ii['ho_saved'], ii['cr_encs']] == list(info[e]), (case, t, e)
            assert bool(done[e]) == dd, (case, t, e)
            err = np.abs(oo - obs[e])
            log.add('obs_elements', ok.size)
            log.add('obs_elements_compared', int(ok.sum()))
            log.err('max_abs_err_obs', err[ok].max())
            assert err[ok].max() <= TOL32, (case, t, e, float(err[ok].max()))
            assert np.abs(rr - rew[e]).max() <= TOL32, (case, t, e)
            post = hw_state_of(arr2, e, Nr, Nc)
            assert post['counter'] == orc.np_random.counter
            assert post['gate_open'] == orc.gate_open and post['bombed'] == orc.bombed
            assert np.array_equal(post['saved'], orc.saved)
            for k in ('rx', 'rv', 'cx', 'cv'):
                d = np.abs(post[k] - getattr(orc, k)).max()
                log.err('max_abs_err_state', d)
                assert d <= TOL32, (case, t, e, k)
            log.add('checked')
        if done.any():
            reset_done(done)
    return log


def hw_golden_teacher_forced(ad, g, args, kw, case):
    Nr, Nh, Nc = args[0], args[1], args[2]
    T = g['actions'].shape[0]
    orc = HostageOracle(*args, rng=Stream(int(g['seed']), int(g['env_id'])), **kw)
    log = ExclusionLog(case, family="hostage", source="reference float64 states (tests/golden)", E=1, T=T,
                       eps=EPS_FRAGILE, tol=TOL32)
    for t in range(T):
        pos = np.concatenate([g['st_rx'][t], g['st_cx'][t], g['st_hx'][t]])
        vel = np.concatenate([g['st_rv'][t], g['st_cv'][t], np.zeros((Nh, 2))])
        flags = 4 | (1 if g['st_gate'][t] else 0) | (2 if g['st_bombed'][t] else 0)
        ad.write(dict(pos_x=pos[None, :, 0], pos_y=pos[None, :, 1], vel_x=vel[None, :, 0], vel_y=vel[None, :, 1],
                      key=g['st_key'][t][None], bomb=g['st_bomb'][t][None],
                      saved=g['st_saved'][t][None].astype(np.uint8), flags=np.array([flags], np.int32),
                      timestep=np.array([g['st_t'][t]], np.int32),
                      rng_counter=np.array([g['st_counter'][t]], np.int64)))
        act = g['actions'][t].astype(np.float32)
        obs, rew, done, info = ad.step(act[None])
        log.add('transitions')
        ref = dict(rx=g['st_rx'][t], rv=g['st_rv'][t], hx=g['st_hx'][t], cx=g['st_cx'][t], cv=g['st_cv'][t],
                   bomb=g['st_bomb'][t], key=g['st_key'][t], saved=g['st_saved'][t],
                   gate_open=bool(g['st_gate'][t]), bombed=bool(g['st_bombed'][t]))
        dyn, ok = hw_fragility(orc, ref, act, EPS_FRAGILE)
        if dyn:
            log.add('excluded_dyn_fragile')
            continue
        assert list(info[0]) == list(g['info'][t]), (case, t)
        assert bool(done[0]) == bool(g['done'][t]), (case, t)
        err = np.abs(g['obs'][t] - obs[0])
        log.add('obs_elements', ok.size)
        log.add('obs_elements_compared', int(ok.sum()))
        log.err('max_abs_err_obs', err[ok].max())
        assert err[ok].max() <= TOL32, (case, t, float(err[ok].max()))
        assert np.abs(g['rew'][t] - rew[0]).max() <= TOL32, (case, t)
        if not g['done'][t]:          # the reference driver reset() after a done step
            post = ad.read()
            assert int(post['rng_counter'][0]) == int(g['st_counter'][t + 1]), (case, t)
            want = np.concatenate([g['st_rx'][t + 1], g['st_cx'][t + 1], g['st_hx'][t + 1]])
            d = max(np.abs(post['pos_x'][0].astype(np.float64) - want[:, 0]).max(),
                    np.abs(post['pos_y'][0].astype(np.float64) - want[:, 1]).max())
            log.err('max_abs_err_state', d)
            assert d <= TOL32, (case, t, d)
            assert np.array_equal(post['saved'][0].astype(bool), g['st_saved'][t + 1])
        log.add('checked')
    return log
