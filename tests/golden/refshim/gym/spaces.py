"""Just enough of gym.spaces for the reference envs (shape/n/nvec/seed/sample)."""
import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = shape
        self.dtype = dtype
        self._rng = np.random.RandomState()

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)
        return [seed]


class Discrete(Space):
    def __init__(self, n):
        super().__init__((), np.int64)
        self.n = int(n)

    def sample(self):
        return int(self._rng.randint(self.n))


class MultiDiscrete(Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        super().__init__(self.nvec.shape, np.int64)

    def sample(self):
        return (self._rng.random_sample(self.nvec.shape) * self.nvec).astype(np.int64)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        super().__init__(tuple(shape) if shape is not None else np.shape(low), dtype)
        self.low = low
        self.high = high

    def sample(self):
        return self._rng.uniform(size=self.shape).astype(self.dtype)


class Dict(Space):
    def __init__(self, spaces):
        super().__init__(None, None)
        self.spaces = spaces

    def seed(self, seed=None):
        for s in self.spaces.values():
            s.seed(seed)
        return [seed]
