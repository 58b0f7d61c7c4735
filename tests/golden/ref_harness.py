"""Runs the UNMODIFIED reference (/root/reference) under the import shim.

Only usable in the build container (the GPU box has no /root/reference); used by
make_golden.py to record golden vectors and by the optional live cross-checks.
Compat patches: np.int alias, int()-coercing Random.randint.
"""
import os
import pickle
import random
import sys

import numpy as np

REFERENCE = os.environ.get("ORLG_REFERENCE", "/root/reference")
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def available():
    return os.path.isdir(os.path.join(REFERENCE, "optical_rl_gym"))


_ready = False


def setup():
    global _ready
    if _ready:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE)
    for p in (REFERENCE, SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not hasattr(np, "int"):
        np.int = int  # removed in numpy 1.24; rwa_env.py:47, rmcsa_env.py:138
    if not getattr(random.Random.randint, "_orlg_patched", False):
        _orig = random.Random.randint

        def randint(self, a, b):  # py3.9 accepted float bounds (rmsa_env.py:40-41)
            return _orig(self, int(a), int(b))

        randint._orlg_patched = True
        random.Random.randint = randint
    import optical_rl_gym  # noqa: F401  (registers the env ids)

    _ready = True


def load_topology(name="nsfnet_chen_5-paths_6-modulations.h5"):
    setup()
    with open(os.path.join(REFERENCE, "examples", "topologies", name), "rb") as f:
        return pickle.load(f)


def make(env_id, **env_args):
    setup()
    import gym

    return gym.make(env_id, **env_args)
