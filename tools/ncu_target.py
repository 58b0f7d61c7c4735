#!/usr/bin/env python
"""Workload for ncu captures: DeepRMSA-v0 NSFNET, N envs, 1000 fill steps + a few more (the profiled ones).
    ncu --set full --clock-control none --import-source on -k regex:deeprmsa_fast -s 1001 -c 2 -o gpurun_out/prof python tools/ncu_target.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-rl-gym_b200"))
import torch  # noqa: E402

from optical_rl_gym_b200 import OpticalVecEnv, nsfnet  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
extra = int(sys.argv[2]) if len(sys.argv) > 2 else 8
env = OpticalVecEnv("DeepRMSA-v0", n, nsfnet(), seed=1, collect_info=False)
a = torch.empty((n, 1), dtype=torch.int32, device="cuda")
for _ in range(1000 + extra):
    env.sample_actions(out=a)
    env.step_raw(a)
torch.cuda.synchronize()
print("done")
