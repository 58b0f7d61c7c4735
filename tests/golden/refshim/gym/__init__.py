"""Minimal stand-in for gym 0.21 so the UNMODIFIED reference package can be
imported in the build container (gym itself is not installable offline).

Test infrastructure only: used by tests/golden/make_golden.py to record golden
vectors from the live reference.  Never imported by the product.
"""
import importlib

import numpy as np

from . import spaces  # noqa: F401
from .envs import registration  # noqa: F401


class Env:
    metadata = {}
    action_space = None
    observation_space = None

    def reset(self):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def render(self, mode="human"):
        return None

    def close(self):
        return None

    def seed(self, seed=None):
        return None

    @property
    def unwrapped(self):
        return self


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.action_space = env.action_space
        self.observation_space = env.observation_space
        self.metadata = getattr(env, "metadata", {})

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        return self.env.step(action)


class ObservationWrapper(Wrapper):
    def reset(self, **kwargs):
        return self.observation(self.env.reset(**kwargs))

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        return self.observation(obs), reward, done, info

    def observation(self, observation):
        raise NotImplementedError


class ActionWrapper(Wrapper):
    def step(self, action):
        return self.env.step(self.action(action))

    def action(self, action):
        raise NotImplementedError


class RewardWrapper(Wrapper):
    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        return obs, self.reward(reward), done, info

    def reward(self, reward):
        raise NotImplementedError


def make(env_id, **kwargs):
    entry = registration.registry[env_id]
    module_name, cls_name = entry.split(":")
    cls = getattr(importlib.import_module(module_name), cls_name)
    return cls(**kwargs)
