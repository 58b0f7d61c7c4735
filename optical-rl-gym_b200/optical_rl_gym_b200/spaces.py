"""Tiny stand-ins for the gym spaces the reference exposes (shape / n / nvec / sample);
gym itself is not a dependency.  (rmsa_env.py:138-149, deeprmsa_env.py:38-43, rmcsa_env.py:181-188)"""
import numpy as np


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.int64

    def __repr__(self):
        return "Discrete(%d)" % self.n


class MultiDiscrete:
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, np.int64)
        self.shape = self.nvec.shape
        self.dtype = np.int64

    def __repr__(self):
        return "MultiDiscrete(%s)" % (self.nvec.tolist(),)


class Box:
    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low, self.high, self.shape, np.dtype(self.dtype).name)


class Dict:
    def __init__(self, spaces):
        self.spaces = dict(spaces)
        self.shape = None

    def __repr__(self):
        return "Dict(%s)" % ", ".join(self.spaces)
