#!/usr/bin/env python
"""orlg_rollout_host with a page-locked result buffer: throughput against the DMA share of the observation rows
(ORLG_HOST_DMA_FRACTION=<share> fixed, ORLG_HOST_DMA=auto adaptive, neither = off).  python tools/time_rollout_dma.py [envs] [steps] [chunk]"""
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
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 2
env = OpticalVecEnv("DeepRMSA-v0", n, nsfnet(), seed=1, collect_info=False, episode_length=1000)
env.rollout(600, "random", want_obs=False, want_actions=False)
ha = torch.from_numpy(np.random.default_rng(0).integers(0, 6, size=(T, n), dtype=np.int32)).pin_memory().numpy()
pinned = os.environ.get("PAGEABLE") is None
ho_t = torch.zeros((T, n, env.obs_dim), dtype=torch.float32)
ho = (ho_t.pin_memory() if pinned else ho_t).numpy()
hr, hd = np.zeros((T, n), np.float32), np.zeros((T, n), np.uint8)
times = []
for rep in range(8):
    t0 = time.perf_counter()
    env.rollout_host(T, "replay", obs=ho, reward=hr, done=hd, actions=ha, chunk=chunk, threads=0)
    times.append(time.perf_counter() - t0)
best = min(times[2:])
print("fraction=%s pinned=%s chunk=%d: %7.1f us per step  %.3e env-steps/s (share now %.3f; calls: %s)" % (
    os.environ.get("ORLG_HOST_DMA_FRACTION"), pinned, chunk, best / T * 1e6, n * T / best, env.host_dma_fraction(),
    " ".join("%.0f" % (t / T * 1e6) for t in times)), flush=True)
