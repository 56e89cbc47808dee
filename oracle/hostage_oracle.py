"""CPU oracle for the ContinuousHostageWorld reset()/step() hot path (TEST INFRASTRUCTURE,
float64 NumPy).

A restatement -- not a copy -- of ``madrl_environments/hostage.py`` (``hw:LINE``); state is held
as plain arrays instead of ``CircAgent`` objects.  Pinned bit-for-bit against the real reference
classes by ``tests/test_hostage_oracle.py`` and the golden vectors in ``tests/golden``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it.
"""
import numpy as np
import scipy.spatial.distance as ssd

from .philox import Stream


class HostageOracle(object):
    """One ContinuousHostageWorld instance; constructor arguments mirror hw:75-79."""

    timestep_limit = 1000  # hw:126-128

    def __init__(self, n_good, n_hostages, n_bad, n_coop_save, n_coop_avoid, radius=0.015,
                 key_loc=None, bad_speed=0.01, n_sensors=30, sensor_range=0.2, action_scale=0.01,
                 save_reward=5., hit_reward=-1., encounter_reward=0.01, not_saved_reward=-3,
                 bomb_reward=-5., bomb_radius=0.05, key_radius=0.0075, control_penalty=-.1,
                 reward_mech='global', addid=True, rng=None):
        self.Nr, self.Nh, self.Nc, self.K = n_good, n_hostages, n_bad, n_sensors
        self.n_coop_save, self.n_coop_avoid = n_coop_save, n_coop_avoid
        self.radius, self.bad_speed = radius, bad_speed
        self.key_loc = None if key_loc is None else np.asarray(key_loc, float).reshape(1, 2)
        self.key_radius, self.bomb_radius = key_radius, bomb_radius
        self.sensor_range, self.action_scale = sensor_range, action_scale
        self.save_reward, self.hit_reward = save_reward, hit_reward
        self.encounter_reward, self.not_saved_reward = encounter_reward, not_saved_reward
        self.bomb_reward, self.control_penalty = bomb_reward, control_penalty
        self.reward_mech, self.addid = reward_mech, addid
        self.r_r, self.r_c, self.r_h = radius, radius, radius * 2      # hw:111-120
        ang = np.linspace(0., 2. * np.pi, n_sensors + 1)[:-1]         # hw:27-29
        self.S = np.c_[np.cos(ang), np.sin(ang)]
        self.obs_dim = n_sensors * 5 + 5 + (1 if addid else 0)        # hw:18-22
        self.np_random = rng if rng is not None else Stream(0, 0)
        self.t = 0
        self.gate_open = False
        self.bombed = False

    # ------------------------------------------------------------------ hw:142-179
    def reset(self):
        rs = self.np_random
        self.t = 0
        self.gate_open = False
        self.bombed = False
        if self.key_loc is None:                     # drawn once, then persists (hw:148-151)
            self.key_loc = 1 - rs.rand(1, 2) * 0.1
        self.rx = np.zeros((self.Nr, 2)); self.rv = np.zeros((self.Nr, 2))
        self.hx = np.zeros((self.Nh, 2))
        self.cx = np.zeros((self.Nc, 2)); self.cv = np.zeros((self.Nc, 2))
        for i in range(self.Nr):
            pos = rs.rand(2)
            pos[-1] = np.clip(pos[-1], 0.55, 0.95)
            self.rx[i] = pos
        for i in range(self.Nh):
            pos = rs.rand(2)
            pos[-1] = np.clip(pos[-1], 0, 0.35 + rs.rand() * 0.01)
            self.hx[i] = pos
        self.saved = np.zeros(self.Nh, dtype=bool)
        for i in range(self.Nc):
            self.cx[i] = rs.rand(2)
            self.cv[i] = rs.rand(2) * self.bad_speed
        self.bomb = np.clip(rs.rand(1, 2), 0., 0.25)
        return self.step(np.zeros((self.Nr, 2)))[0]

    @property
    def is_terminal(self):                            # hw:181-184
        return bool(self.bombed or self.saved.all() or self.t >= self.timestep_limit)

    def _sensed(self, i, objx, same=False):           # hw:62-70
        rel = objx - self.rx[i][None, :]
        sv = self.S.dot(rel.T)
        bad = (sv < 0) | (sv > self.sensor_range) | (
            (rel ** 2).sum(axis=1)[None, :] - sv ** 2 > self.r_r ** 2)
        sv[bad] = np.inf
        if same:
            sv[:, i] = np.inf
        return sv

    @staticmethod
    def _caught(coll, n_coop):                        # hw:186-199
        caught = np.where(coll.sum(axis=0) >= n_coop)[0]
        who = np.where(coll[:, caught] >= 1)[0]
        return caught, who

    def _closest(self, sv_all):                       # hw:201-207, 319-351
        idx = np.argmin(sv_all, axis=2)
        d = np.take_along_axis(sv_all, idx[:, :, None], axis=2)[:, :, 0]
        mask = np.isfinite(d)
        return np.where(mask, d, 0.0), idx, mask

    # ------------------------------------------------------------------ hw:228-429
    def step(self, action):
        Nr = self.Nr
        act = np.asarray(action, dtype=np.float64).reshape((Nr, 2)) * self.action_scale
        rewards = np.zeros(Nr)
        self.rv = self.rv + act                                               # hw:236-238
        self.rx = self.rx + self.rv
        if self.reward_mech == 'global':                                      # hw:241-244
            rewards += self.control_penalty * (act ** 2).sum()
        else:
            rewards += self.control_penalty * (act ** 2).sum(axis=1)
        clipped = np.clip(self.rx, 0, 1)                                      # hw:247-252
        self.rv[self.rx != clipped] = 0
        self.rx = clipped
        if not self.gate_open:                                                # hw:255-261
            clipped = np.clip(self.rx, 0.5 + self.radius, 1)
            self.rv[self.rx != clipped] *= -1
            self.rx = clipped
        # collisions hw:264-296
        coll_ho = ssd.cdist(self.rx, self.hx) <= self.r_r + self.r_h
        ho_caught, who_ho = self._caught(coll_ho, self.n_coop_save)
        ho_enc, who_ho_enc = self._caught(coll_ho, 1)
        coll_cr = ssd.cdist(self.rx, self.cx) <= self.r_r + self.r_c
        cr_caught, who_cr = self._caught(coll_cr, 1)
        coll_bo = ssd.cdist(self.rx, self.bomb) <= self.r_r + self.bomb_radius
        bo_caught, who_bo = self._caught(coll_bo, 1)
        coll_ke = ssd.cdist(self.rx, self.key_loc) <= self.r_r + self.key_radius
        ke_caught, _ = self._caught(coll_ke, 1)
        # sensing hw:300-315 (saved hostages masked with the PRE-step mask)
        sv_ho = np.array([self._sensed(i, self.hx) for i in range(Nr)])
        sv_ho[:, :, self.saved] = np.inf
        sv_cr = np.array([self._sensed(i, self.cx) for i in range(Nr)])
        sv_bo = np.array([self._sensed(i, self.bomb) for i in range(Nr)])
        sv_ke = np.array([self._sensed(i, self.key_loc) for i in range(Nr)])
        f_ho, _, _ = self._closest(sv_ho)
        if not self.gate_open:                                                # hw:323-325
            f_ho = np.zeros_like(f_ho)
        f_cr, idx_cr, m_cr = self._closest(sv_cr)
        f_bo, _, _ = self._closest(sv_bo)
        f_ke, _, _ = self._closest(sv_ke)
        if self.gate_open:                                                    # hw:343-345
            f_ke = np.zeros_like(f_ke)
        s_cr = np.zeros((Nr, self.K))                                         # hw:359-361
        for i in range(Nr):
            rv = self.S.dot((self.cv - self.rv[i][None, :]).T)
            s_cr[i] = np.where(m_cr[i], rv[np.arange(self.K), idx_cr[i]], 0.0)
        # process collisions hw:368-381
        rs = self.np_random
        self.saved[ho_caught] = True
        for j in cr_caught:
            self.cx[j] = rs.rand(2)
            self.cv[j] = (rs.rand(2) - 0.5) * self.bad_speed
        if bo_caught.size:
            self.bombed = True
        if ke_caught.size:
            self.gate_open = True
        # rewards hw:383-392 (post-update gate / bomb flags)
        if self.reward_mech == 'global':
            rewards += (len(ho_enc) * self.encounter_reward * self.gate_open +
                        len(ho_caught) * self.save_reward + len(cr_caught) * self.hit_reward +
                        self.bombed * self.bomb_reward)
        else:
            rewards[who_ho] += self.save_reward
            rewards[who_ho_enc] += self.encounter_reward * self.gate_open
            rewards[who_cr] += self.hit_reward
            rewards[who_bo] += self.bombed * self.bomb_reward
        feats = np.c_[f_cr, s_cr, f_ho, f_ke, f_bo]                           # hw:395-397
        # criminals drift hw:399-404
        self.cx = self.cx + self.cv
        flip = np.all(self.cx != np.clip(self.cx, 0, 1), axis=1)
        self.cv[flip] = -1 * self.cv[flip]
        obs = []                                                              # hw:406-421
        for i in range(Nr):
            tail = [float(coll_ho[i].sum() > 0), float(coll_cr[i].sum() > 0),
                    float(coll_ke[i].sum() > 0), float(coll_bo[i].sum() > 0), float(self.gate_open)]
            if self.addid:
                tail.append(i + 1)
            obs.append(np.concatenate([feats[i], tail]))
        self.t += 1                                                           # hw:423-427
        done = self.is_terminal
        if done:
            rewards += np.sum(~self.saved) * self.not_saved_reward
        return obs, rewards, done, dict(ho_saved=len(ho_caught), cr_encs=len(cr_caught))

    # ------------------------------------------------------------------ state access (tests)
    def get_state(self):
        return dict(rx=self.rx.copy(), rv=self.rv.copy(), hx=self.hx.copy(), cx=self.cx.copy(),
                    cv=self.cv.copy(), bomb=self.bomb.copy(), key=self.key_loc.copy(),
                    saved=self.saved.copy(), gate_open=self.gate_open, bombed=self.bombed, t=self.t,
                    counter=getattr(self.np_random, 'counter', 0))

    def set_state(self, s):
        for k in ('rx', 'rv', 'hx', 'cx', 'cv', 'bomb'):
            setattr(self, k, np.array(s[k], dtype=np.float64))
        self.key_loc = np.array(s['key'], dtype=np.float64).reshape(1, 2)
        self.saved = np.array(s['saved'], dtype=bool)
        self.gate_open, self.bombed, self.t = bool(s['gate_open']), bool(s['bombed']), int(s['t'])
        if 'counter' in s and hasattr(self.np_random, 'counter'):
            self.np_random.counter = int(s['counter'])
