"""Per-predicate fragility analysis for the float paths (TEST INFRASTRUCTURE, float64 NumPy).

The reference computes in float64; the production kernels in float32.  A float32 result can only be
held to 1e-5 of the float64 one where no comparison of the step sits within rounding distance of
its threshold.  Round 1 skipped a whole transition when ANY of its (up to 72 600) predicates was
near a threshold; this module instead classifies each predicate:

  * `dyn`  -- predicates that change the trajectory (wall clips, obstacle rebound, collisions,
    evader / criminal bounce, gate clip): if one of them is within `eps`, the whole transition is
    excluded (they are few -- Np + Nobj + Np*Nobj -- and almost never near);
  * `obs_ok[Np, D]` -- one flag per observation element: a sensor feature (distance and speed of
    class c at sensor k of agent i) is excluded only if one of ITS in/out-of-beam tests
    (ww:64-72 / hw:62-70) or the argmin between ITS two best candidates is within `eps`.

Only ``tests/`` and ``__graft_entry__.smoke()`` import this.
"""
import numpy as np
import scipy.spatial.distance as ssd


def _near(a, b, eps):
    d = np.abs(np.asarray(a, dtype=np.float64) - b)
    return bool(np.any((d < eps) & (d > 0)))


def _bounce_near(x2, eps):
    """ww:397-409 flips the velocity iff BOTH coordinates are outside [0,1]: fragile only if one
    coordinate is within eps of a wall while the other is outside (or itself near)."""
    x2 = np.asarray(x2, float)
    near = ((np.abs(x2) < eps) | (np.abs(x2 - 1.0) < eps)) & (x2 != 0.0) & (x2 != 1.0)
    out = (x2 < 0) | (x2 > 1) | near
    return bool(np.any((near[:, 0] & out[:, 1]) | (near[:, 1] & out[:, 0])))


def sensor_fragile(S, rel, r2, rng, eps, exclude=None, dead=None):
    """[K] bool: would the nearest-object feature of a sensor change discretely (another object
    selected, or sensed <-> not sensed) if every compared quantity moved by up to eps?
    S [K,2] sensor unit vectors, rel [N,2] object positions relative to the sensing agent."""
    sv = S.dot(rel.T)
    q = (rel ** 2).sum(axis=1)[None, :] - sv ** 2
    lo = (sv < -eps) | (sv > rng + eps) | (q > r2 + eps)
    hi = (sv < eps) | (sv > rng - eps) | (q > r2 - eps)
    if exclude is not None:
        lo[:, exclude] = hi[:, exclude] = True
    if dead is not None:
        lo[:, dead] = hi[:, dead] = True
    frag = (lo != hi).any(axis=1)
    if sv.shape[1] >= 2:
        part = np.sort(np.where(hi, np.inf, sv), axis=1)
        two = np.isfinite(part[:, 1])
        with np.errstate(invalid='ignore'):
            frag |= two & (part[:, 1] - part[:, 0] < eps)
    return frag


def ww_fragility(o, state, action, eps):
    """(dyn_fragile, obs_ok[Np, D]) for one MAWaterWorld step of oracle `o` from `state`."""
    px = np.asarray(state['px'], float); pv = np.asarray(state['pv'], float)
    ex = np.asarray(state['ex'], float); ev = np.asarray(state['ev'], float)
    ox = np.asarray(state['ox'], float); ov = np.asarray(state['ov'], float)
    obst = np.asarray(state['obst'], float).reshape(1, 2)
    K, Np = o.K, o.Np
    obs_ok = np.ones((Np, o.obs_dim), dtype=bool)
    act = np.asarray(action, float).reshape(Np, 2) * o.action_scale
    pv = pv + act
    px = px + pv
    dyn = _near(px, 0.0, eps) or _near(px, 1.0, eps)
    px = np.clip(px, 0, 1)
    dyn = dyn or _near(ssd.cdist(px, obst), o.r_p + o.obstacle_radius, eps)
    dyn = dyn or _near(ssd.cdist(ex, obst), o.r_e + o.obstacle_radius, eps)
    dyn = dyn or _near(ssd.cdist(ox, obst), o.r_po + o.obstacle_radius, eps)
    dyn = dyn or _near(ssd.cdist(px, ex), o.r_p + o.r_e, eps)
    dyn = dyn or _near(ssd.cdist(px, ox), o.r_p + o.r_po, eps)
    dyn = dyn or _bounce_near(ex + ev, eps) or _bounce_near(ox + ov, eps)
    # feature columns per class (ww:388-395): speed -> ob, ev, evs, po, pos, pu, pus; else ob, ev, po, pu
    if o.speed_features:
        cols = {'ob': [0], 'ev': [1, 2], 'po': [3, 4], 'pu': [5, 6]}
    else:
        cols = {'ob': [0], 'ev': [1], 'po': [2], 'pu': [3]}
    for name, objx, same in (('ob', obst, False), ('ev', ex, False), ('po', ox, False), ('pu', px, True)):
        for i in range(Np):
            frag = sensor_fragile(o.S, objx - px[i][None, :], o.r_p ** 2, o.sensor_range, eps,
                                  exclude=i if same else None)
            for c in cols[name]:
                obs_ok[i, c * K:(c + 1) * K] &= ~frag
    return bool(dyn), obs_ok


def hw_fragility(o, state, action, eps):
    """(dyn_fragile, obs_ok[Nr, D]) for one ContinuousHostageWorld step (obs layout hw:395-421:
    crdist, crspeed, hodist, kedist, bodist, then the tail)."""
    rx = np.asarray(state['rx'], float); rv = np.asarray(state['rv'], float)
    hx = np.asarray(state['hx'], float); cx = np.asarray(state['cx'], float)
    cv = np.asarray(state['cv'], float)
    bomb = np.asarray(state['bomb'], float).reshape(1, 2)
    key = np.asarray(state['key'], float).reshape(1, 2)
    saved = np.asarray(state['saved'], bool)
    K, Nr = o.K, o.Nr
    obs_ok = np.ones((Nr, o.obs_dim), dtype=bool)
    act = np.asarray(action, float).reshape(Nr, 2) * o.action_scale
    rv = rv + act
    rx = rx + rv
    dyn = _near(rx, 0.0, eps) or _near(rx, 1.0, eps)
    rx = np.clip(rx, 0, 1)
    if not state['gate_open']:
        dyn = dyn or _near(rx, 0.5 + o.radius, eps)
        rx = np.clip(rx, 0.5 + o.radius, 1)
    dyn = dyn or _near(ssd.cdist(rx, hx), o.r_r + o.r_h, eps)
    dyn = dyn or _near(ssd.cdist(rx, cx), o.r_r + o.r_c, eps)
    dyn = dyn or _near(ssd.cdist(rx, bomb), o.r_r + o.bomb_radius, eps)
    dyn = dyn or _near(ssd.cdist(rx, key), o.r_r + o.key_radius, eps)
    dyn = dyn or _bounce_near(cx + cv, eps)
    for cols, objx, dead, live in (([0, 1], cx, None, True), ([2], hx, saved, bool(state['gate_open'])),
                                   ([3], key, None, not state['gate_open']), ([4], bomb, None, True)):
        if not live:            # feature forced to zero this step (hw:323-325, 343-345)
            continue
        for i in range(Nr):
            frag = sensor_fragile(o.S, objx - rx[i][None, :], o.r_r ** 2, o.sensor_range, eps, dead=dead)
            for c in cols:
                obs_ok[i, c * K:(c + 1) * K] &= ~frag
    return bool(dyn), obs_ok


class ExclusionLog(object):
    """Counts what a teacher-forced comparison checked and what it excluded, per test case; the
    GPU tests dump it as JSON (gpurun_out/parity/*.json on the box; a copy is committed under
    profiles/)."""

    def __init__(self, case, **meta):
        self.d = dict(case=case, transitions=0, checked=0, excluded_dyn_fragile=0, excluded_other=0,
                      obs_elements=0, obs_elements_compared=0, max_abs_err_obs=0.0,
                      max_abs_err_state=0.0, **meta)

    def add(self, key, n=1):
        self.d[key] += n

    def err(self, key, v):
        self.d[key] = max(self.d[key], float(v))

    @property
    def checked_frac(self):
        return self.d['checked'] / max(1, self.d['transitions'])

    @property
    def obs_frac(self):
        return self.d['obs_elements_compared'] / max(1, self.d['obs_elements'])

    def dump(self, root):
        import json
        import os
        out = os.path.join(root, "gpurun_out", "parity")
        try:
            os.makedirs(out, exist_ok=True)
            d = dict(self.d, checked_frac=self.checked_frac, obs_compared_frac=self.obs_frac)
            with open(os.path.join(out, self.d['case'] + ".json"), "w") as f:
                json.dump(d, f, indent=1, sort_keys=True)
        except OSError:
            pass
        return self.d
