"""Minimal observation/action space descriptors.

The reference uses ``gym.spaces.Box`` / ``Discrete`` (``madrl_environments/__init__.py:2``); gym
is not a dependency of this package, so equivalent light-weight classes are provided.  If gym is
importable its classes are used instead so that ``isinstance`` checks in adapters such as
``rllabwrapper/__init__.py:16-27`` (``convert_gym_space``) keep working.
"""
import numpy as np

try:  # pragma: no cover - gym is absent from the build image
    from gym.spaces import Box, Discrete  # type: ignore
except Exception:

    class Box(object):
        """Box(low, high, shape=None): same construction rules as gym.spaces.Box."""

        def __init__(self, low, high, shape=None):
            if shape is None:
                self.low = np.asarray(low, dtype=np.float64)
                self.high = np.asarray(high, dtype=np.float64)
                assert self.low.shape == self.high.shape
            else:
                self.low = np.full(shape, low, dtype=np.float64)
                self.high = np.full(shape, high, dtype=np.float64)

        @property
        def shape(self):
            return self.low.shape

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool((x >= self.low).all() and (x <= self.high).all())

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return np.random.uniform(lo, hi)

        def __repr__(self):
            return "Box" + str(self.shape)

        def __eq__(self, other):
            return isinstance(other, Box) and np.array_equal(self.low, other.low) and \
                np.array_equal(self.high, other.high)

    class Discrete(object):
        """Discrete(n): {0, ..., n-1}."""

        def __init__(self, n):
            self.n = n

        @property
        def shape(self):
            return ()

        def contains(self, x):
            return int(x) == x and 0 <= int(x) < self.n

        def sample(self):
            return np.random.randint(self.n)

        def __repr__(self):
            return "Discrete(%d)" % self.n

        def __eq__(self, other):
            return isinstance(other, Discrete) and self.n == other.n
