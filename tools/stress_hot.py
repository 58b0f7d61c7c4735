#!/usr/bin/env python
"""Long-run stress of the rollout path at BASELINE size: N DeepRMSA envs x T steps (Philox traffic, random policy, HOT
kernel instance, dependent launches), then invariants over all envs and a bit-for-bit comparison of sampled envs with
the oracle (which replays the same Philox streams).   python tools/stress_hot.py [envs] [steps] [samples]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "optical-rl-gym_b200"))
import torch  # noqa: E402

from optical_rl_gym_b200 import OpticalVecEnv, nsfnet  # noqa: E402
from oracle import oracle  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 96
seed = 99
tables = nsfnet()
env = OpticalVecEnv("DeepRMSA-v0", n, tables, seed=seed, episode_length=777)
a = torch.empty((n, 1), dtype=torch.int32, device="cuda")
for t in range(T):
    env.sample_actions(out=a)
    env.step_raw(a)
torch.cuda.synchronize()
assert int(env.error_flags().abs().sum()) == 0
cnt = env.counters().cpu().numpy()
assert (cnt[:, 0] == T + 1).all()
m, alloc, now, nheap = env.export_state(allocation=True)
avail = env.available_slots()
assert torch.equal((avail == 0), (alloc.reshape(avail.shape) >= 0)), "busy slots != slots owned by live services"
avail = avail.cpu().numpy()
obs = env.observation().cpu().numpy()
rng = np.random.default_rng(0)
idx = sorted(set([0, 1, 31, 32, n - 1] + rng.integers(0, n, ns).tolist()))
for i in idx:
    o = oracle.OracleEnv("DeepRMSA-v0", tables, num_slots=100, episode_length=777)
    o.set_philox(seed, i)
    o.reset(full=True)
    o.rollout(T, policy=1)
    oa, oal, onow, onh = o.state()
    assert np.array_equal(avail[i].reshape(oa.shape), oa), ("masks", i)
    assert np.array_equal(alloc[i].cpu().numpy(), oal), ("allocation", i)
    assert now[i].item() == onow and nheap[i].item() == onh, ("clock / live services", i)
    assert np.array_equal(cnt[i], o.counters()), ("counters", i)
    np.testing.assert_allclose(obs[i], o.observation(), rtol=1e-6, atol=0)
print("stress ok: %d envs x %d steps, %d envs compared with the oracle, accept rate %.4f, max live services %d"
      % (n, T, len(idx), cnt[:, 1].sum() / cnt[:, 0].sum(), int(nheap.max())))
