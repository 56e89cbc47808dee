"""Counter-based random stream shared by the oracle and the CUDA engine (TEST INFRASTRUCTURE).

The reference draws its randomness from NumPy's MT19937 (``self.np_random.rand`` in
waterworld.py:141-170,360-374 and hostage.py:149-177,374-375; global ``np.random.randint`` /
``np.random.uniform`` in pursuit_evade.py:183-186 and utils/agent_utils.py:39-45; the evader
controller ``rng.randint`` in utils/Controllers.py:16).  A sequential Mersenne twister cannot be
reproduced per-env on a GPU, so parity is defined on *injected identical streams* (SURVEY.md 8c):
both the reference/oracle (through the objects below) and the CUDA kernels
(``madrl_b200/csrc/philox.cuh``) consume this stream, in the reference's own draw order.

Stream definition (the spec -- the .cuh twin must agree bit for bit):

  * generator: Philox4x32-10 (Salmon et al., SC'11), multipliers 0xD2511F53 / 0xCD9E8D57,
    Weyl constants 0x9E3779B9 / 0xBB67AE85;
  * key      = (seed & 0xffffffff, seed >> 32);
  * counter  = (block & 0xffffffff, block >> 32, env_id, tag), block = draw_index // 4;
  * draw n of env ``env_id`` is word ``n % 4`` of that block; ``tag`` separates stream families
    (0 = environment dynamics, 1 = benchmark action generator);
  * u32 -> [0,1):  (u >> 8) * 2**-24   (exact in fp32 and fp64);
  * u32 -> {lo..hi-1}:  lo + ((u * (hi - lo)) >> 32)   (multiply-shift, integer only).
"""
import numpy as np

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """One Philox4x32-10 block: 4 counter words + 2 key words -> 4 output words (python ints)."""
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def u32_to_unit(u):
    return (u >> 8) * (1.0 / 16777216.0)


def u32_to_range(u, lo, hi):
    return lo + ((u * (hi - lo)) >> 32)


class Stream(object):
    """Sequential view of one env's counter-based stream.

    Exposes the subset of the ``numpy.random.RandomState`` interface that the reference envs call
    (``rand``, ``randint``, ``uniform``, ``random_sample``) so that it can be injected as
    ``env.np_random`` / a controller ``rng`` / a stand-in for the ``np.random`` module.
    ``counter`` is the number of u32 draws consumed so far (the engine keeps the same number in
    its per-env state).
    """

    def __init__(self, seed=0, env_id=0, tag=0, counter=0):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.env_id = int(env_id) & MASK
        self.tag = int(tag) & MASK
        self.counter = int(counter)
        self._blk = None
        self._words = None

    def next_u32(self):
        blk = self.counter >> 2
        if blk != self._blk:
            self._words = philox4x32_10(blk & MASK, (blk >> 32) & MASK, self.env_id, self.tag,
                                        self.seed & MASK, (self.seed >> 32) & MASK)
            self._blk = blk
        w = self._words[self.counter & 3]
        self.counter += 1
        return w

    # --- numpy.random.RandomState look-alikes -------------------------------------------------
    def random_sample(self, size=None):
        if size is None:
            return u32_to_unit(self.next_u32())
        out = np.empty(size, dtype=np.float64)
        flat = out.reshape(-1)
        for i in range(flat.size):
            flat[i] = u32_to_unit(self.next_u32())
        return out

    def rand(self, *shape):
        if len(shape) == 0:
            return u32_to_unit(self.next_u32())
        return self.random_sample(shape)

    def randint(self, low, high=None, size=None):
        if high is None:
            low, high = 0, low
        assert size is None
        return u32_to_range(self.next_u32(), int(low), int(high))

    def uniform(self, low=0.0, high=1.0, size=None):
        assert size is None
        # numpy: low + (high - low) * random_sample()
        return low + (high - low) * u32_to_unit(self.next_u32())


class StreamController(object):
    """Evader/pursuer controller drawing its actions from a Stream
    (drop-in for utils/Controllers.py:8-16 ``RandomPolicy``)."""

    def __init__(self, n_actions, stream):
        self.n_actions = n_actions
        self.rng = stream

    def act(self, state):
        return self.rng.randint(self.n_actions)


def philox_block_numpy(block, env_id, tag, seed):
    """Vectorised Philox block for arrays of (block, env_id): returns uint32 array [..., 4]."""
    block = np.asarray(block, dtype=np.uint64)
    env_id = np.asarray(env_id, dtype=np.uint64)
    block, env_id = np.broadcast_arrays(block, env_id)
    c0 = (block & MASK).astype(np.uint64)
    c1 = (block >> np.uint64(32)).astype(np.uint64)
    c2 = env_id.astype(np.uint64) & np.uint64(MASK)
    c3 = np.full(c0.shape, tag, dtype=np.uint64)
    k0 = np.uint64(seed & MASK)
    k1 = np.uint64((seed >> 32) & MASK)
    m = np.uint64(MASK)
    s32 = np.uint64(32)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        c0, c1, c2, c3 = ((p1 >> s32) ^ c1 ^ k0) & m, p1 & m, ((p0 >> s32) ^ c3 ^ k1) & m, p0 & m
        k0 = (k0 + np.uint64(W0)) & m
        k1 = (k1 + np.uint64(W1)) & m
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)
