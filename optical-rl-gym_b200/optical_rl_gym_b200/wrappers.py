"""Batched versions of the reference's gym wrappers over :class:`OpticalVecEnv`.

Same class names and constructor arguments as the reference (``rmsa_env.py:806-874``,
``rwa_env.py:505-536``, ``rmcsa_env.py:914-947``, ``wrappers.py:4-16``); the wrapped object keeps the
VecEnv protocol (``reset`` / ``step`` / ``step_async`` / ``step_wait`` / attribute pass-through) and all
tensors stay on the device.
"""
from __future__ import annotations

import numpy as np
import torch

from . import spaces


class VecWrapper:
    def __init__(self, venv):
        self.venv = venv
        self.action_space = venv.action_space
        self.observation_space = venv.observation_space

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.venv, name)

    @property
    def unwrapped(self):
        return getattr(self.venv, "unwrapped", self.venv)

    def reset(self, **kwargs):
        return self.venv.reset(**kwargs)

    def step_async(self, actions):
        self.venv.step_async(actions)

    def step_wait(self):
        return self.venv.step_wait()

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()


class SimpleMatrixObservation(VecWrapper):
    """Observation = one-hot(min(src, dst)) ++ one-hot(max(src, dst)) ++ ``available_slots`` flattened.
    The reference declares ``Box(uint8)`` but returns float64; here ``dtype`` picks the tensor type
    (uint8 by default, ``torch.float64`` for the reference's values)."""

    def __init__(self, venv, dtype=torch.uint8):
        super().__init__(venv)
        base = self.unwrapped
        shape = 2 * base.tables.num_nodes + base.tables.num_links * base.num_spectrum_resources * base.num_spatial_resources
        self.observation_space = spaces.Box(0, 1, (shape,), np.uint8)
        self._dtype = dtype
        self._buf = None

    def observation(self, observation=None):
        self._buf = self.unwrapped.matrix_observation(out=self._buf)
        return self._buf if self._dtype == torch.uint8 else self._buf.to(self._dtype)

    def reset(self, **kwargs):
        self.venv.reset(**kwargs)
        return self.observation()

    def step_wait(self):
        _, reward, done, info = self.venv.step_wait()
        return self.observation(), reward, done, info


class PathOnlyFirstFitAction(VecWrapper):
    """The agent selects the path only; the slot / wavelength is the first fit on that path
    (RMSA: ``range(0, S - n)`` like the reference, i.e. never the last feasible start)."""

    def __init__(self, venv):
        super().__init__(venv)
        base = self.unwrapped
        if base.env_id not in ("RMSA-v0", "RWA-v0"):
            raise NotImplementedError("PathOnlyFirstFitAction: RMSA-v0 / RWA-v0 (the RMCSA one raises in the reference)")
        self.action_space = spaces.Discrete(base.k_paths + base.reject_action)
        self._mapped = None

    def action(self, action):
        self._mapped = self.unwrapped.path_only_first_fit(action, out=self._mapped)
        return self._mapped

    def step_async(self, actions):
        self.venv.step_async(self.action(actions))


class UseInfoReward(VecWrapper):
    """reward = info[info_key] (``wrappers.py:4-16``)."""

    def __init__(self, venv, info_key):
        super().__init__(venv)
        self.info_key = info_key

    def reward(self, reward, info):
        return info[self.info_key]

    def step_wait(self):
        obs, reward, done, info = self.venv.step_wait()
        return obs, self.reward(reward, info), done, info
