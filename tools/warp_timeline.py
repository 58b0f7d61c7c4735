#!/usr/bin/env python
"""Distribution of per-warp entry/exit times of one fast-kernel launch (instrumented build)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-rl-gym_b200"))
import torch
from optical_rl_gym_b200 import OpticalVecEnv, _native, nsfnet
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = OpticalVecEnv("DeepRMSA-v0", n, nsfnet(), seed=1, collect_info=False)
a = torch.empty((n, 1), dtype=torch.int32, device="cuda")
for _ in range(1200):
    env.sample_actions(out=a); env.step_raw(a)
torch.cuda.synchronize()
L = _native.lib()
nw = n // 32
buf = (C.c_ulonglong * (2 * nw))()
assert L.orlg_debug_warp_timeline(buf, nw) == 0
t = np.frombuffer(buf, dtype=np.uint64).reshape(nw, 2).astype(np.int64)
t0 = t[:, 0].min()
start, end = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3
dur = end - start
pct = lambda x: " ".join("%.1f" % v for v in np.percentile(x, [0, 10, 50, 90, 99, 100]))
print("warps", nw, "kernel span %.1f us" % end.max())
print("entry  us p0/10/50/90/99/100:", pct(start))
print("exit   us p0/10/50/90/99/100:", pct(end))
print("dur    us p0/10/50/90/99/100:", pct(dur))
