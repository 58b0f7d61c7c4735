#!/usr/bin/env python
"""orlg_rollout_host (host actions in, float32 rows out) against the decoder thread count and the pipeline chunk.
    python tools/time_rollout_host.py [envs] [steps]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-rl-gym_b200"))
from optical_rl_gym_b200 import OpticalVecEnv, nsfnet  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 40
env = OpticalVecEnv("DeepRMSA-v0", n, nsfnet(), seed=1, collect_info=False, episode_length=1000)
env.rollout(600, "random", want_obs=False, want_actions=False)
ha = torch.from_numpy(np.random.default_rng(0).integers(0, 6, size=(T, n), dtype=np.int32)).pin_memory().numpy()
ho, hr, hd = np.zeros((T, n, env.obs_dim), np.float32), np.zeros((T, n), np.float32), np.zeros((T, n), np.uint8)
cores = os.cpu_count() or 1
for policy in ("replay", "random"):
    for chunk in (2, 4, 8):
        for nt in sorted({cores, cores - 1, cores - 2}):
            if nt < 1:
                continue
            best = 1e9
            for _ in range(4):
                t0 = time.perf_counter()
                env.rollout_host(T, policy, obs=ho, reward=hr, done=hd, actions=ha if policy == "replay" else None, chunk=chunk, threads=nt)
                best = min(best, time.perf_counter() - t0)
            print("%-6s chunk %d threads %2d: %7.1f us per step  %.3g env-steps/s" % (policy, chunk, nt, best / T * 1e6, n * T / best), flush=True)
