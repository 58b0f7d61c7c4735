#!/usr/bin/env python
"""Per-warp wall-clock timeline of one orlg_rollout launch (needs a build with -DORLG_RO_TIMELINE):
    tools/build_variant.sh tl -DORLG_RO_TIMELINE ; python tools/rollout_timeline.py [envs] [T]
Shows where a T-step launch spends its time: entry (state in), the steps, window rebuilds (how many warps take one,
how long it lasts), exit, and how far the last warp trails the median one."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-rl-gym_b200"))
import torch  # noqa: E402

from optical_rl_gym_b200 import OpticalVecEnv, _native, nsfnet  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
env = OpticalVecEnv("DeepRMSA-v0", n, nsfnet(), seed=1, collect_info=False)
obs = torch.empty((T, n, env.obs_dim), dtype=torch.float32, device="cuda")
for _ in range(1000 // T + 3):
    env.rollout(T, obs=obs)
torch.cuda.synchronize()
L = _native.lib()
nw = (n + 31) // 32
pct = lambda x: " ".join("%7.1f" % v for v in np.percentile(x, [0, 10, 50, 90, 99, 100]))
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); env.rollout(T, obs=obs); b.record()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * (8 * nw))()
    assert L.orlg_debug_rollout_timeline(buf, nw) == 0, L.orlg_last_error()
    t = np.frombuffer(buf, dtype=np.uint64).reshape(nw, 8).astype(np.int64)
    t0 = t[:, 0].min()
    us = lambda col: (t[:, col] - t0) / 1e3
    print("launch %d: %.1f us by CUDA events; warps %d; last warp exits at %.1f us" % (rep, a.elapsed_time(b) * 1e3, nw, us(3).max()))
    print("  kernel entry          us p0/10/50/90/99/100:", pct(us(0)))
    print("  first step starts     us p0/10/50/90/99/100:", pct(us(1)))
    print("  last step ends        us p0/10/50/90/99/100:", pct(us(2)))
    print("  exit                  us p0/10/50/90/99/100:", pct(us(3)))
    steps = (t[:, 2] - t[:, 1]) / 1e3
    rb = t[:, 5] / 1e3
    print("  step loop             us p0/10/50/90/99/100:", pct(steps))
    print("  step loop - rebuilds  us p0/10/50/90/99/100:", pct(steps - rb))
    for k in range(0, 4):
        m = t[:, 4] == k
        if m.any():
            print("  warps with %d rebuild(s): %5d  loop %.1f us (median), rebuild time %.1f us each" % (
                k, int(m.sum()), float(np.median(steps[m])), float(np.median(rb[m] / max(k, 1)))))
    m = t[:, 4] == 1
    if m.any():
        print("  inside a rebuild (warps with one): scan %.1f us, sort %.1f us, rest (flush, head loads, pops) %.1f us" % (
            float(np.median(t[m, 6])) / 1e3, float(np.median(t[m, 7])) / 1e3, float(np.median(t[m, 5] - t[m, 6] - t[m, 7])) / 1e3))
