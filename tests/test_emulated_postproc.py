"""The trajectory post-processing kernels (madrl_b200/csrc/postproc.cu: GAE, frame stack,
standardiser, episode statistics, whole-batch moments / advantage centring) executed on the CPU by
the warp emulator of tests/emu and compared with oracle/postproc_oracle.py -- the GPU test
tests/test_postproc.py::test_cuda_* at small sizes, through the same C ABI."""
import ctypes as C
import os
import shutil

import numpy as np
import pytest

from oracle import postproc_oracle as po

pytestmark = pytest.mark.skipif(shutil.which(os.environ.get("CXX", "g++")) is None, reason="no host C++ compiler")


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


@pytest.fixture(scope="module")
def lib():
    from emu import driver
    L = driver.load(())
    vp, i32, dbl = C.c_void_p, C.c_int, C.c_double
    L.madrl_gae_f32.argtypes = [i32, i32, i32, vp, vp, vp, vp, dbl, dbl, vp, vp, vp]
    L.madrl_frame_stack_f32.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    L.madrl_standardize_f32.argtypes = [i32, C.c_size_t, vp, vp, vp, dbl, dbl, i32, dbl, i32, vp]
    L.madrl_episode_stats_f32.argtypes = [i32, i32, i32, vp, vp, dbl, i32, vp, vp, vp, vp, vp, vp]
    L.madrl_moments_f32.argtypes = [C.c_size_t, vp, vp, vp, vp, vp]
    L.madrl_center_advantages_f32.argtypes = [C.c_size_t, vp, i32, i32, vp, vp, vp]
    return L


def _data():
    rs = np.random.RandomState(2)
    T, E, A, D, B = 30, 5, 3, 7, 4
    return dict(T=T, E=E, A=A, D=D, B=B, obs0=rs.randn(E, A, D).astype(np.float32),
                obs=rs.randn(T, E, A, D).astype(np.float32), rew=rs.randn(T, E, A).astype(np.float32),
                val=rs.randn(T, E, A).astype(np.float32), last=rs.randn(E, A).astype(np.float32),
                done=(rs.rand(T, E) < 0.12).astype(np.uint8))


def test_gae(lib):
    d = _data()
    T, E, A = d['T'], d['E'], d['A']
    for lv in (None, d['last']):
        adv, ret = np.empty_like(d['rew']), np.empty_like(d['rew'])
        assert lib.madrl_gae_f32(T, E, A, _p(d['rew']), _p(d['val']), _p(d['done']), _p(lv), 0.99, 0.95, _p(adv),
                                 _p(ret), None) == 0
        for e in range(E):
            a_o, r_o = po.gae_env(d['rew'][:, e], d['val'][:, e], d['done'][:, e].astype(bool), 0.99, 0.95,
                                  None if lv is None else lv[e])
            assert np.abs(adv[:, e] - a_o).max() < 1e-4 and np.abs(ret[:, e] - r_o).max() < 1e-4


def test_frame_stack(lib):
    d = _data()
    T, E, A, D, B = d['T'], d['E'], d['A'], d['D'], d['B']
    carry = np.repeat(d['obs0'][..., None], B, axis=-1).copy()          # reset(): every slot = reset obs
    out = np.empty((T, E, A, D, B), np.float32)
    for t0, t1 in ((0, 11), (11, T)):                                       # two calls exercise the carry
        o = np.ascontiguousarray(d['obs'][t0:t1])
        dn = np.ascontiguousarray(d['done'][t0:t1])
        res = np.empty((t1 - t0, E, A, D, B), np.float32)
        assert lib.madrl_frame_stack_f32(t1 - t0, E, A, D, B, _p(o), _p(dn), _p(carry), _p(res), None) == 0
        out[t0:t1] = res
    for e in range(E):
        s0, so = po.frame_stack_env(d['obs0'][e].astype(np.float64), d['obs'][:, e].astype(np.float64),
                                    d['done'][:, e].astype(bool), B)
        assert np.array_equal(out[:, e], so.astype(np.float32))


def test_standardizer(lib):
    d = _data()
    T, E, A, D = d['T'], d['E'], d['A'], d['D']
    n = E * A * D
    mean, var = np.zeros(n), np.ones(n)
    x0, xo = d['obs0'].copy(), d['obs'].copy()
    assert lib.madrl_standardize_f32(1, n, _p(x0), _p(mean), _p(var), 0.05, 1e-8, 1, 1.0, 1, None) == 0
    assert lib.madrl_standardize_f32(T, n, _p(xo), _p(mean), _p(var), 0.05, 1e-8, 1, 1.0, 1, None) == 0
    rmean, rvar = np.zeros(E * A), np.ones(E * A)
    xr = d['rew'].copy()
    assert lib.madrl_standardize_f32(T, E * A, _p(xr), _p(rmean), _p(rvar), 0.02, 1e-8, 0, 0.5, 1, None) == 0
    for e in range(E):
        m = po.StandardizeEnv(A, D, 0.5, True, True, 0.05, 0.02)
        assert np.abs(m.obs(d['obs0'][e].astype(np.float64)) - x0[e]).max() < 1e-5
        for t in range(T):
            assert np.abs(m.obs(d['obs'][t, e].astype(np.float64)) - xo[t, e]).max() < 1e-5
            assert np.abs(m.rew(d['rew'][t, e].astype(np.float64)) - xr[t, e]).max() < 1e-5
        assert np.abs(m.obs_var - var.reshape(E, A, D)[e]).max() < 1e-12


def test_episode_stats(lib):
    d = _data()
    T, E, A = d['T'], d['E'], d['A']
    carry = np.zeros((E, A + 3))
    outs = dict(r=np.empty((T, E, A), np.float32), dsc=np.empty((T, E), np.float32),
                ln=np.empty((T, E), np.int32), end=np.empty((T, E), np.uint8))
    for t0, t1 in ((0, 13), (13, T)):
        r = np.ascontiguousarray(d['rew'][t0:t1]); dn = np.ascontiguousarray(d['done'][t0:t1])
        part = [np.empty((t1 - t0, E, A), np.float32), np.empty((t1 - t0, E), np.float32),
                np.empty((t1 - t0, E), np.int32), np.empty((t1 - t0, E), np.uint8)]
        assert lib.madrl_episode_stats_f32(t1 - t0, E, A, _p(r), _p(dn), 0.9, 12, _p(carry), *[_p(x) for x in part],
                                           None) == 0
        for k, x in zip(('r', 'dsc', 'ln', 'end'), part):
            outs[k][t0:t1] = x
    for e in range(E):
        recs = po.episode_stats_env(d['rew'][:, e].astype(np.float64), d['done'][:, e].astype(bool), 0.9, 12)
        assert [r[0] for r in recs] == list(np.nonzero(outs['end'][:, e])[0])
        for t, ep_r, avg, disc, length in recs:
            assert np.abs(outs['r'][t, e] - ep_r).max() < 1e-4 and abs(outs['dsc'][t, e] - disc) < 1e-4
            assert outs['ln'][t, e] == length


def test_moments_and_advantage_centring(lib):
    rs = np.random.RandomState(9)
    for n in (1, 3, 1050, 20000):        # 20000: every block of the (narrowed, see build_emu.py) grid loops
        adv = (rs.randn(n) * 3 + 1).astype(np.float32)
        ret = rs.randn(n).astype(np.float32)
        stats, ws = np.empty(9), np.empty(4096)
        assert lib.madrl_moments_f32(n, _p(adv), _p(ret), _p(stats), _p(ws), None) == 0
        a64, r64 = adv.astype(np.float64), ret.astype(np.float64)
        want = np.array([a64.mean(), r64.mean(), (r64 - a64).mean(), a64.var(), r64.var(), (r64 - a64).var(),
                         a64.min(), r64.min(), (r64 - a64).min()])
        assert np.allclose(stats, want, rtol=1e-10, atol=1e-12)
        for center, positive in ((1, 0), (0, 1), (1, 1)):
            x = adv.copy()
            assert lib.madrl_center_advantages_f32(n, _p(x), center, positive, _p(stats), _p(ws), None) == 0
            w = a64
            if center:
                w = po.center_advantages(w)
            if positive:
                w = po.shift_advantages_to_positive(w)
            assert np.abs(x - w).max() <= 1e-5 * max(1.0, np.abs(w).max())


def test_path_packing_kernels_equal_host_to_paths(lib):
    """madrl_paths_plan + madrl_paths_pack_u32 (the device-side `to_paths`) against the host triple
    loop of madrl_b200.postproc.to_paths: same paths in the same order, with and without the one-step
    observation shift."""
    from madrl_b200.postproc import to_paths
    vp, i32 = C.c_void_p, C.c_int
    lib.madrl_paths_plan.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.madrl_paths_pack_u32.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    rs = np.random.RandomState(4)
    for T, E, A, D, p_done in ((30, 5, 3, 7, 0.12), (9, 2, 1, 40, 0.4), (17, 3, 4, 1, 0.0)):
        obs = rs.randn(T, E, A, D).astype(np.float32)
        obs_before = rs.randn(E, A, D).astype(np.float32)
        act = rs.randint(0, 5, size=(T, E, A)).astype(np.int32)
        rew = rs.randn(T, E, A).astype(np.float32)
        info = rs.randint(0, 9, size=(T, E)).astype(np.int32)
        done = (rs.rand(T, E) < p_done).astype(np.uint8)
        seg_s, seg_l = np.full((T, E), -1, np.int32), np.full((T, E), -1, np.int32)
        n_ep, ep_rec = np.zeros(E, np.int32), np.zeros((E, T, 3), np.int32)
        assert lib.madrl_paths_plan(T, E, A, _p(done), _p(seg_s), _p(seg_l), _p(n_ep), _p(ep_rec), None) == 0

        def pack(x, n_agents, first=None, per_agent=True):
            tail = x.shape[3:] if per_agent else x.shape[2:]
            out = np.zeros((T * E * n_agents,) + tail, x.dtype)
            Dw = max(1, int(np.prod(tail)))
            assert lib.madrl_paths_pack_u32(T, E, n_agents, Dw, _p(np.ascontiguousarray(x)), _p(first), _p(seg_s),
                                            _p(seg_l), _p(out), None) == 0
            return out
        for shifted in (False, True):
            p_obs = pack(obs, A, obs_before if shifted else None)
            p_act, p_rew, p_info = pack(act, A), pack(rew, A), pack(info, 1, per_agent=False)
            src_obs = np.concatenate([obs_before[None], obs[:-1]]) if shifted else obs
            want = to_paths(src_obs, act, rew, done, dict(k=info))
            got = []
            for e in range(E):
                for j in range(n_ep[e]):
                    s, L, term = ep_rec[e, j]
                    for a in range(A):
                        o = e * T * A + s * A + a * L
                        got.append(dict(observations=p_obs[o:o + L], actions=p_act[o:o + L], rewards=p_rew[o:o + L],
                                        env_infos=dict(k=p_info[e * T + s:e * T + s + L]), env=e, agent=a,
                                        terminated=bool(term)))
            assert len(got) == len(want)
            for g, w in zip(got, want):
                assert (g['env'], g['agent'], g['terminated']) == (w['env'], w['agent'], w['terminated'])
                for k in ('observations', 'actions', 'rewards'):
                    assert np.array_equal(g[k], w[k]), k
                assert np.array_equal(g['env_infos']['k'], w['env_infos']['k'])


def test_standardizer_with_terminal_observations_follows_the_reference_order(lib):
    """madrl_standardize_obs_terminal_f32 vs the StandardizedEnv oracle (pinned to the real wrapper in
    tests/test_postproc.py): at a done step the estimate sees the terminal observation first, then the
    reset observation (step() then reset(), madrl_environments/__init__.py:283-291)."""
    lib.madrl_standardize_obs_terminal_f32.argtypes = [C.c_int, C.c_int, C.c_size_t] + [C.c_void_p] * 5 + \
        [C.c_double, C.c_double, C.c_void_p]
    d = _data()
    T, E, A, D = d['T'], d['E'], d['A'], d['D']
    rs = np.random.RandomState(8)
    term = rs.randn(T, E, A, D).astype(np.float32)
    x, tm = d['obs'].copy(), term.copy()
    mean, var = np.zeros((E, A, D)), np.ones((E, A, D))
    assert lib.madrl_standardize_obs_terminal_f32(T, E, A * D, _p(x), _p(tm), _p(d['done']), _p(mean), _p(var),
                                                  0.05, 1e-8, None) == 0
    for e in range(E):
        mine = po.StandardizeEnv(A, D, 1.0, True, False, 0.05, 0.001)
        for t in range(T):
            if d['done'][t, e]:
                assert np.abs(np.array(mine.obs(term[t, e].astype(np.float64))) - tm[t, e]).max() < 1e-5, (t, e)
            else:
                assert np.array_equal(tm[t, e], term[t, e])          # untouched
            assert np.abs(np.array(mine.obs(d['obs'][t, e].astype(np.float64))) - x[t, e]).max() < 1e-5, (t, e)
        assert np.abs(np.array(mine.obs_mean) - mean[e]).max() < 1e-12
