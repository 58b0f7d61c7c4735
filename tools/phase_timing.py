#!/usr/bin/env python
"""Per-phase cycle breakdown of the fast DeepRMSA kernel (needs a build with -DORLG_PHASE_TIMING):
    ORLG_NVCC_EXTRA=-DORLG_PHASE_TIMING python optical-rl-gym_b200/optical_rl_gym_b200/build.py
    python tools/phase_timing.py [envs] [steps]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-rl-gym_b200"))
import torch  # noqa: E402

from optical_rl_gym_b200 import OpticalVecEnv, _native, nsfnet  # noqa: E402

NAMES = ["issue loads", "wait tables+sync", "A: decision+push", "B: traffic draw", "wait masks", "alloc+releases",
         "path AND", "dirty writeback", "barrier 1", "features+obs row", "scalar stores", "barrier 2", "tile copy-out"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
env = OpticalVecEnv("DeepRMSA-v0", n, nsfnet(), seed=1, collect_info=False)
a = torch.empty((n, 1), dtype=torch.int32, device="cuda")
for _ in range(1000):
    env.sample_actions(out=a)
    env.step_raw(a)
torch.cuda.synchronize()
L = _native.lib()
buf = (C.c_ulonglong * 16)()
L.orlg_debug_phase_cycles(buf)
for _ in range(steps):
    env.sample_actions(out=a)
    env.step_raw(a)
torch.cuda.synchronize()
assert L.orlg_debug_phase_cycles(buf) == 0, L.orlg_last_error()
warps = (n + 31) // 32
tot = sum(buf[:13])
print("avg cycles per warp per step: %.0f" % (tot / warps / steps))
for i, nm in enumerate(NAMES):
    print("%-20s %8.0f cycles  %5.1f%%" % (nm, buf[i] / warps / steps, 100.0 * buf[i] / tot))
for i, nm in ((13, "  rel: directory"), (14, "  rel: group scan"), (15, "  rel: payload+apply")):
    print("%-20s %8.0f cycles" % (nm, buf[i] / warps / steps))
