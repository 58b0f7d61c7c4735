#!/usr/bin/env python
"""Per-phase cycle breakdown of the rollout kernel (needs a build with -DORLG_PHASE_TIMING):
    ORLG_NVCC_EXTRA=-DORLG_PHASE_TIMING python optical-rl-gym_b200/optical_rl_gym_b200/build.py
    python tools/rollout_phases.py [envs] [T] [launches]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-rl-gym_b200"))
import torch  # noqa: E402

from optical_rl_gym_b200 import OpticalVecEnv, _native, nsfnet  # noqa: E402

NAMES = ["request draw", "action + phase A", "phase B + releases", "window rebuild", "tile acquire", "features",
         "obs rows -> tile", "tile copy-out + release", "ENTRY (per launch)", "EXIT (per launch)", "path AND"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 5
env = OpticalVecEnv("DeepRMSA-v0", n, nsfnet(), seed=1, collect_info=False)
obs = torch.empty((T, n, env.obs_dim), dtype=torch.float32, device="cuda")
for _ in range(1000 // T + 1):
    env.rollout(T, obs=obs)
torch.cuda.synchronize()
L = _native.lib()
buf = (C.c_ulonglong * 16)()
L.orlg_debug_phase_cycles(buf)
for _ in range(launches):
    env.rollout(T, obs=obs)
torch.cuda.synchronize()
assert L.orlg_debug_phase_cycles(buf) == 0, L.orlg_last_error()
warps = (n + 31) // 32
steps = T * launches
tot = sum(buf[:8]) + buf[10]
print("avg cycles per warp per step: %.0f (+ entry %.0f, exit %.0f per launch); warp rebuilds per step: %.4f" % (
    tot / warps / steps, buf[8] / warps / launches, buf[9] / warps / launches, buf[15] / warps / steps))
for i, nm in [(j, NAMES[j]) for j in (0, 1, 2, 3, 10, 5, 4, 6, 7)]:
    print("%-22s %8.0f cycles  %5.1f%%" % (nm, buf[i] / warps / steps, 100.0 * buf[i] / tot))
for i, nm in ((11, "  rebuild: flush"), (12, "  rebuild: scan"), (13, "  rebuild: sort")):
    print("%-22s %8.0f cycles per rebuild" % (nm, buf[i] / max(buf[15], 1)))
