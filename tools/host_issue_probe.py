#!/usr/bin/env python
"""Host-side issue rate of the rollout loop (two launches per step) against the GPU time per step:
    python tools/host_issue_probe.py      (on a B200)"""
import sys
import time
sys.path.insert(0, "optical-rl-gym_b200")
import torch
from optical_rl_gym_b200 import OpticalVecEnv, nsfnet
n = 65536
env = OpticalVecEnv("DeepRMSA-v0", n, nsfnet(), seed=1, collect_info=False)
a = torch.empty((n, 1), dtype=torch.int32, device="cuda")
for _ in range(300):
    env.sample_actions(out=a); env.step_raw(a)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(200):
        env.sample_actions(out=a); env.step_raw(a)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("host issue: %.2f us/step   wall incl. drain: %.2f us/step" % ((t1 - t0) / 200 * 1e6, (t2 - t0) / 200 * 1e6))
