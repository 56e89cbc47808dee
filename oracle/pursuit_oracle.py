"""CPU oracle for the PursuitEvade reset()/step() hot path (TEST INFRASTRUCTURE, NumPy).

A restatement -- not a copy -- of ``madrl_environments/pursuit/pursuit_evade.py`` (``pe:LINE``)
and its helpers ``pursuit/utils/DiscreteAgent.py`` (``da:``), ``AgentLayer.py`` (``al:``),
``agent_utils.py`` (``au:``), ``Controllers.py`` (``ct:``).  Agents are rows of integer arrays
with "gone" flags instead of lists of objects that get popped; the random draws (map choice,
constraint window, rejection-sampled spawns, evader actions) all come from one injected stream
in the reference's order.  Pinned bit-for-bit against the real reference by
``tests/test_pursuit_oracle.py`` and the golden vectors in ``tests/golden``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it.
"""
import numpy as np

from .philox import Stream

# da:28-38  0 left, 1 right, 2 up, 3 down, 4 stay
MOTION = np.array([[-1, 0], [1, 0], [0, 1], [0, -1], [0, 0]], dtype=np.int64)
SURROUND = np.array([[-1, 0], [1, 0], [0, 1], [0, -1]], dtype=np.int64)  # pe:150


class PursuitOracle(object):
    """One PursuitEvade instance (train_pursuit=True).  kwargs mirror pe:49-148."""

    def __init__(self, map_pool, n_evaders=1, n_pursuers=1, obs_range=3, flatten=True,
                 layer_norm=10, n_catch=2, catchr=0.01, term_pursuit=5.0, urgency_reward=0.0,
                 include_id=True, surround=True, constraint_window=1.0, sample_maps=False,
                 reward_mech='global', random_opponents=False, max_opponents=10, rng=None):
        self.map_pool = np.asarray(map_pool)
        self.map = self.map_pool[0]
        self.xs, self.ys = self.map.shape
        self.Ne, self.Np, self.R = n_evaders, n_pursuers, obs_range
        self.off = int((obs_range - 1) / 2)                                   # pe:65
        self.flatten, self.layer_norm, self.n_catch = flatten, layer_norm, n_catch
        self.catchr, self.term_pursuit, self.urgency_reward = catchr, term_pursuit, urgency_reward
        self.include_id, self.surround = include_id, surround
        self.constraint_window, self.sample_maps = constraint_window, sample_maps
        self.reward_mech = reward_mech
        self.random_opponents, self.max_opponents = random_opponents, max_opponents       # pe:81-82
        self.rng = rng if rng is not None else Stream(0, 0)
        self.local_obs = np.zeros((n_pursuers, 4, obs_range, obs_range))        # pe:119, never re-zeroed
        self.model_state = np.zeros((4,) + self.map.shape, dtype=np.float32)    # pe:152
        self.ppos = np.zeros((n_pursuers, 2), dtype=np.int64)                   # au:22 agents start at (0,0)
        self.epos = np.zeros((n_evaders, 2), dtype=np.int64)
        self.gone = np.zeros(n_evaders, dtype=bool)

    # ------------------------------------------------------------------ helpers
    def _count_grid(self, pos, live=None):                                      # al:50-64
        g = np.zeros((self.xs, self.ys), dtype=np.int32)
        for i in range(len(pos)):
            if live is None or live[i]:
                g[pos[i, 0], pos[i, 1]] += 1
        return g

    def _move(self, pos, a):                                                    # da:69-97
        if self.map[pos[0], pos[1]] == -1:       # in a building: frozen (never happens after reset)
            return
        x, y = pos[0] + MOTION[a, 0], pos[1] + MOTION[a, 1]
        if not (0 <= x < self.xs and 0 <= y < self.ys):
            return
        if self.map[x, y] == -1:
            return
        pos[0], pos[1] = x, y

    def _spawn(self, xl, xu, yl, yu):                                           # au:31-47
        while True:
            x = self.rng.randint(xl, xu)
            y = self.rng.randint(yl, yu)
            if self.map[x, y] != -1:
                return x, y

    # ------------------------------------------------------------------ pe:173-207
    def reset(self):
        self.gone[:] = False
        n_ev = self.Ne
        if self.random_opponents:            # pe:177-181: this episode has 1 .. max_opponents-1 evaders;
            n_ev = self.rng.randint(1, self.max_opponents)   # the draw precedes the map sample
            self.gone[n_ev:] = True          # the others never exist (the reference rebuilds its lists)
        if self.sample_maps:
            self.map = self.map_pool[self.rng.randint(len(self.map_pool))]
        xws = self.rng.uniform(0.0, 1.0 - self.constraint_window)
        yws = self.rng.uniform(0.0, 1.0 - self.constraint_window)
        xl, xu = int(self.xs * xws), int(self.xs * (xws + self.constraint_window))
        yl, yu = int(self.ys * yws), int(self.ys * (yws + self.constraint_window))
        for i in range(self.Np):
            self.ppos[i] = self._spawn(xl, xu, yl, yu)
        for i in range(n_ev):
            self.epos[i] = self._spawn(xl, xu, yl, yu)
        self.model_state[0] = self.map
        self.model_state[1] = self._count_grid(self.ppos)
        self.model_state[2] = self._count_grid(self.epos, ~self.gone)
        return self._collect_obs()

    # ------------------------------------------------------------------ pe:359-381
    def _reward(self):
        es = self._count_grid(self.epos, ~self.gone)
        r = np.zeros(self.Np)
        for i in range(self.Np):
            xx = np.clip(self.ppos[i, 0] + SURROUND[:, 0], 0, self.xs - 1)
            yy = np.clip(self.ppos[i, 1] + SURROUND[:, 1], 0, self.ys - 1)
            r[i] = self.catchr * np.sum(es[xx, yy])
        return r

    # ------------------------------------------------------------------ pe:523-540
    def _need_to_surround(self, x, y):
        need = 4
        if x == 0 or x == self.xs - 1:
            need -= 1
        if y == 0 or y == self.ys - 1:
            need -= 1
        for d in SURROUND:
            xn, yn = x + d[0], y + d[1]
            if not 0 < xn < self.xs or not 0 < yn < self.ys:   # NB: row/column 0 is skipped too
                continue
            if self.map[xn, yn] == -1:
                need -= 1
        return need

    # ------------------------------------------------------------------ pe:463-521
    def _remove(self):
        pgrid = self.model_state[1]
        sur = np.zeros(self.Np, dtype=bool)
        removed = 0
        newly = []
        for i in range(self.Ne):
            if self.gone[i]:
                continue
            x, y = self.epos[i]
            if self.surround:
                adj = []                      # neighbour cells holding >= 1 pursuer
                for d in SURROUND:
                    xn, yn = x + d[0], y + d[1]
                    if 0 <= xn < self.xs and 0 <= yn < self.ys and pgrid[xn, yn] > 0:
                        adj.append((xn, yn))
                if len(adj) == self._need_to_surround(x, y):
                    newly.append(i)
                    removed += 1
                    for j in range(self.Np):
                        if (self.ppos[j, 0], self.ppos[j, 1]) in adj:
                            sur[j] = True
            else:
                if pgrid[x, y] >= self.n_catch:
                    newly.append(i)
                    removed += 1
                    for j in range(self.Np):
                        if self.ppos[j, 0] == x and self.ppos[j, 1] == y:
                            sur[j] = True
        for i in newly:
            self.gone[i] = True
        return removed, sur

    # ------------------------------------------------------------------ pe:418-461
    def _collect_obs(self):
        obs = []
        for i in range(self.Np):
            lo = self.local_obs[i]
            lo[0].fill(1.0 / self.layer_norm)
            x, y = self.ppos[i]
            xlo, xhi = max(x - self.off, 0), min(x + self.off, self.xs - 1)
            ylo, yhi = max(y - self.off, 0), min(y + self.off, self.ys - 1)
            xo, yo = xlo - (x - self.off), ylo - (y - self.off)
            lo[0:3, xo:xo + (xhi - xlo) + 1, yo:yo + (yhi - ylo) + 1] = \
                np.abs(self.model_state[0:3, xlo:xhi + 1, ylo:yhi + 1]) / self.layer_norm
            lo[3, self.R // 2, self.R // 2] = float(i) / self.Np
            if self.flatten:
                o = lo[0:3].flatten()
                if self.include_id:
                    o = np.append(o, float(i) / self.Np)
                obs.append(o)
            else:
                obs.append(np.rollaxis(lo, 0, 3).copy())
        return obs

    # ------------------------------------------------------------------ pe:209-262
    def step(self, actions):
        rewards = self._reward()                        # from the PRE-move state
        if isinstance(actions, (list, np.ndarray)):
            acts = list(actions)
        else:                                           # joint scalar action, pe:233
            acts = list(np.unravel_index(actions, [5] * self.Np))
        for i, a in enumerate(acts):
            self._move(self.ppos[i], int(a))
        for i in range(self.Ne):                        # live evaders, in order: one draw each
            if not self.gone[i]:
                self._move(self.epos[i], self.rng.randint(5))
        self.model_state[0] = self.map
        self.model_state[1] = self._count_grid(self.ppos)
        self.model_state[2] = self._count_grid(self.epos, ~self.gone)
        removed, sur = self._remove()                   # model_state[2] is NOT refreshed: a
        obs = self._collect_obs()                       # captured evader is still visible now
        rewards += self.term_pursuit * sur
        rewards += self.urgency_reward
        done = bool(self.gone.all())                    # pe:384-389
        if self.reward_mech == 'global':
            return obs, [rewards.mean()] * self.Np, done, {'removed': removed}
        return obs, rewards, done, {'removed': removed}

    # ------------------------------------------------------------------ state access (tests)
    def get_state(self):
        return dict(ppos=self.ppos.copy(), epos=self.epos.copy(), gone=self.gone.copy(),
                    local_obs=self.local_obs.copy(), map=self.map.copy(),
                    counter=getattr(self.rng, 'counter', 0))
