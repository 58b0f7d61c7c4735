"""Stand-ins for the gym spaces the reference exposes (gym itself is not a dependency): shape / n / nvec / dtype,
``seed()``, ``sample()`` and ``contains()`` with gym 0.21 semantics, which is what Stable-Baselines3's VecEnv consumers
(`action_space.sample()` for warm-up / random agents, shape and dtype checks) use.
(rmsa_env.py:138-149, deeprmsa_env.py:38-43, rwa_env.py:60-72, rmcsa_env.py:181-188)"""
import numpy as np


class Space:
    shape = None
    dtype = None

    def __init__(self):
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        """Re-seeds this space's sampler (gym: returns the list of seeds used)."""
        self._rng = np.random.default_rng(seed)
        return [seed]

    def __contains__(self, x):
        return self.contains(x)


class Discrete(Space):
    def __init__(self, n):
        super().__init__()
        self.n = int(n)
        self.shape = ()
        self.dtype = np.int64

    def sample(self):
        return int(self._rng.integers(self.n))

    def contains(self, x):
        try:
            v = int(x)
        except (TypeError, ValueError):
            return False
        return np.ndim(x) == 0 and v == x and 0 <= v < self.n

    def __repr__(self):
        return "Discrete(%d)" % self.n


class MultiDiscrete(Space):
    def __init__(self, nvec):
        super().__init__()
        self.nvec = np.asarray(nvec, np.int64)
        self.shape = self.nvec.shape
        self.dtype = np.int64

    def sample(self):
        return (self._rng.random(self.nvec.shape) * self.nvec).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and np.issubdtype(x.dtype, np.integer) and bool(((x >= 0) & (x < self.nvec)).all())

    def __repr__(self):
        return "MultiDiscrete(%s)" % (self.nvec.tolist(),)


class Box(Space):
    def __init__(self, low, high, shape, dtype):
        super().__init__()
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def sample(self):
        lo = np.broadcast_to(np.asarray(self.low, np.float64), self.shape)
        hi = np.broadcast_to(np.asarray(self.high, np.float64), self.shape)
        if np.issubdtype(np.dtype(self.dtype), np.integer):
            return self._rng.integers(lo.astype(np.int64), hi.astype(np.int64) + 1).astype(self.dtype)
        return self._rng.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool((x >= self.low).all() and (x <= self.high).all())

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low, self.high, self.shape, np.dtype(self.dtype).name)


class Dict(Space):
    def __init__(self, spaces):
        super().__init__()
        self.spaces = dict(spaces)
        self.shape = None

    def seed(self, seed=None):
        for i, sp in enumerate(self.spaces.values()):
            sp.seed(None if seed is None else seed + i)
        return [seed]

    def sample(self):
        return {k: sp.sample() for k, sp in self.spaces.items()}

    def contains(self, x):
        return isinstance(x, dict) and set(x) == set(self.spaces) and all(sp.contains(x[k]) for k, sp in self.spaces.items())

    def __repr__(self):
        return "Dict(%s)" % ", ".join(self.spaces)
