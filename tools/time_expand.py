#!/usr/bin/env python
"""Host decoder (orlg_expand_packed) alone: ns per row and GB/s delivered against the number of host threads, on packed records
of a real DeepRMSA rollout (GPU needed to produce them).    python tools/time_expand.py [envs] [steps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-rl-gym_b200"))
from optical_rl_gym_b200 import OpticalVecEnv, nsfnet  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
env = OpticalVecEnv("DeepRMSA-v0", n, nsfnet(), seed=1, collect_info=False, episode_length=1000)
env.rollout(600, "random", want_obs=False, want_actions=False)
pk = env.rollout_packed(T, "random").cpu().numpy()
rows = n * T
obs = np.zeros((T, n, env.obs_dim), np.float32)
import ctypes as C
lib = env._lib
rew, done, act = np.zeros((T, n), np.float32), np.zeros((T, n), np.uint8), np.zeros((T, n), np.int32)
for nt in (1, 2, 4, 8, 12, 16, 24, 32):
    if nt > 2 * (os.cpu_count() or 1):
        break
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        rc = lib.orlg_expand_packed(pk.ctypes.data_as(C.c_void_p), rows, 14, 100, obs.ctypes.data_as(C.c_void_p), rew.ctypes.data_as(C.c_void_p),
                                    done.ctypes.data_as(C.c_void_p), act.ctypes.data_as(C.c_void_p), nt)
        best = min(best, time.perf_counter() - t0)
    assert rc == 0
    print("%2d threads: %6.1f ns/row/thread  %6.1f GB/s delivered  %7.1f us per %d rows  -> %.3g env-steps/s" % (
        nt, best / rows * 1e9 * nt, rows * 221 / best / 1e9, best / T * 1e6, n, rows / best))
