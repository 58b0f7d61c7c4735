"""Times orlg_rollout (persistent kernel) against the per-step path on the headline workload.
usage: python tools/time_rollout.py [envs] [T] [reps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "optical-rl-gym_b200")):
    sys.path.insert(0, p)
import torch

from optical_rl_gym_b200 import OpticalVecEnv, nsfnet

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
policy = os.environ.get("POLICY", "random")
kind = os.environ.get("KIND", "DeepRMSA-v0")
kargs = {"DeepRMSA-v0": {}, "RMSA-v0": dict(load=250, mean_service_holding_time=25), "RWA-v0": dict(load=450, mean_service_holding_time=25)}[kind]
env = OpticalVecEnv(kind, n, nsfnet(), seed=1, collect_info=False, episode_length=1000, **kargs)
obs = torch.empty((T, n, env.obs_dim), dtype=torch.float32, device="cuda") if env.obs_dim else None
rew = torch.empty((T, n), dtype=torch.float32, device="cuda")
done = torch.empty((T, n), dtype=torch.uint8, device="cuda")
act = torch.empty((T, n, env.action_dim), dtype=torch.int32, device="cuda")
fill = 1000
for _ in range(fill // T + 1):
    env.rollout(T, policy, obs=obs, reward=rew, done=done, actions=act)
torch.cuda.synchronize()
times = []
for r in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    env.rollout(T, policy, obs=obs, reward=rew, done=done, actions=act)
    b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
times.sort()
try:
    import pynvml
    pynvml.nvmlInit()
    hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
    clk = "sm %d MHz (max %d), %.0f W" % (pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetMaxClockInfo(hnd, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(hnd) / 1e3)
except Exception as exc:  # noqa: BLE001
    clk = "clocks n/a (%s)" % exc
med = times[len(times) // 2]
print(kind, policy, "rollout: n=%d T=%d span=%s warps=%s tiles=%s  median %.3f ms/launch = %.2f us/step, %.3e env-steps/s (min %.3f max %.3f)  accept %.3f err %d" % (
    n, T, os.environ.get("ORLG_RO_SPAN"), os.environ.get("ORLG_RO_WARPS"), os.environ.get("ORLG_RO_TILES"),
    med, med / T * 1e3, n * T / (med * 1e-3), times[0], times[-1], float((rew > 0).float().mean()),
    int((env.error_flags() != 0).sum())), clk)
if os.environ.get("STEP_PATH"):
    a1 = torch.empty((n, env.action_dim), dtype=torch.int32, device="cuda")
    pol_fn = (lambda: env.sample_actions(out=a1)) if policy == "random" else (lambda: env.heuristic(policy, out=a1))
    for _ in range(50):
        pol_fn(); env.step_raw(a1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(500):
        pol_fn(); env.step_raw(a1)
    b.record()
    torch.cuda.synchronize()
    print("step path: %.2f us/step" % (a.elapsed_time(b) / 500 * 1e3))
