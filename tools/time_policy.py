"""PPO-policy variant of the headline config: obs -> MLP policy -> env.step at 65536 envs.
Times (CUDA events) the fused tensor-core policy kernel alone, the torch float32 / bf16 forward, and the full
per-step loop with either policy."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "optical-rl-gym_b200")):
    sys.path.insert(0, p)
import torch

from optical_rl_gym_b200 import OpticalVecEnv, nsfnet
from optical_rl_gym_b200.policy import MlpPolicy

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
torch.manual_seed(0)
pol = MlpPolicy(54, 5, (128,) * 5).cuda()
for m in pol.shared_net:
    if isinstance(m, torch.nn.Linear):
        torch.nn.init.orthogonal_(m.weight, gain=2 ** 0.5)
env = OpticalVecEnv("DeepRMSA-v0", n, nsfnet(), seed=1, collect_info=False, episode_length=1000)
env.rollout(1000, want_obs=False, want_actions=False)
obs = env.observation().contiguous()
act = torch.empty((n, 1), dtype=torch.int32, device="cuda")


def timed(fn, reps=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(400000)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


flop = 2.0 * n * (54 * 128 + 4 * 128 * 128 + 6 * 128)
t = timed(lambda: pol.act_native(obs, out=act))
print("fused policy kernel (tcgen05, bf16): %.2f us per %d envs = %.1f TFLOP/s" % (t, n, flop / t / 1e6))
t32 = timed(lambda: pol.act(obs), reps=50)
print("torch float32 forward + argmax:      %.2f us" % t32)
pol16 = MlpPolicy(54, 5, (128,) * 5).cuda().to(torch.bfloat16)
t16 = timed(lambda: pol16.act(obs.to(torch.bfloat16)), reps=50)
print("torch bf16 forward + argmax:         %.2f us" % t16)


def step_native():
    pol.act_native(env._obs, out=act)
    env.step_raw(act)


def step_torch():
    env.step_raw(pol.act(env._obs))


env.observation()
ts = timed(step_native, reps=500)
print("per step, fused policy + step kernel: %.2f us = %.3e env-steps/s" % (ts, n / ts * 1e6))
tt = timed(step_torch, reps=100)
print("per step, torch f32 policy + step:    %.2f us = %.3e env-steps/s" % (tt, n / tt * 1e6))
print("errors:", int((env.error_flags() != 0).sum()))
