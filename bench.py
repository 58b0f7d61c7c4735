#!/usr/bin/env python
"""Benchmark of the batched step hot path (BASELINE.json metric: env-steps/s, DeepRMSA-v0 NSFNET,
65536 envs per GPU, obs + reward + done on device).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one VecEnv.step of the whole batch (one env-step for each of the 65536 envs of every
rank) with the uniform random policy drawn on the device.  Prints ONE JSON line (rank 0).
See DESIGN.md "Measurement" for the definitions of value / e2e / roofline / cpu_baseline.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "optical-rl-gym_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

ENV_ID = "DeepRMSA-v0"
ENV_ARGS = dict(episode_length=1000)          # DeepRMSA defaults: k=5, j=1, 100 slots, holding 25 / iat 0.1 = 250 Erlang
ENVS_PER_GPU = 65536
FILL_STEPS = 1000                              # ~4 mean holding times: network in steady state before timing
# algorithmic bytes per env-step for this config (random policy, fp32 obs)
B_STEP = 868
WORKLOAD = "DeepRMSA-v0 NSFNET(14n/22l) k=5 j=1 S=100 250E, %d envs/GPU, uniform random policy, Philox traffic"


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback"


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.01)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def oracle_kwargs():
    # stats=False: like the GPU arm, the CPU arm computes obs / reward / done / counters, not the float statistics of info
    return dict(num_slots=100, episode_length=ENV_ARGS["episode_length"], j=1, mean_holding=25.0,
                mean_iat=1 / float((25.0 / 0.1) / 25.0), stats=False)


def cpu_baseline(threads, seconds=12.0):
    """The oracle port on all host cores, on a bounded sample of the same workload."""
    from optical_rl_gym_b200 import nsfnet
    from oracle import oracle

    n_envs = max(64, 16 * threads)
    vec = oracle.OracleVec(ENV_ID, nsfnet(), n_envs, seed=1, threads=threads, **oracle_kwargs())
    vec.run(FILL_STEPS)                                     # steady state, untimed
    chunk, done_steps, t0 = 500, 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        vec.run(chunk)
        done_steps += chunk
    dt = time.perf_counter() - t0
    vec.close()
    return {"value": n_envs * done_steps / dt, "unit": "env-steps/s", "cores": threads, "kind": "port",
            "sample": "%d envs x %d steps after a %d-step fill, oracle/orlg_oracle.c with float64 observation, "
                      "%d pthreads" % (n_envs, done_steps, FILL_STEPS, threads)}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference is
    pure Python and cannot travel to the GPU box), all host threads, same config / metric / unit."""
    if rank != 0:
        return
    from optical_rl_gym_b200 import nsfnet
    from oracle import oracle

    threads = os.cpu_count() or 1
    n_envs = int(os.environ.get("ORLG_REF_ENVS", 8192))
    vec = oracle.OracleVec(ENV_ID, nsfnet(), n_envs, seed=1, threads=threads, **oracle_kwargs())
    vec.run(FILL_STEPS)
    if args.warmup:
        vec.run(args.warmup)
    t0 = time.perf_counter()
    vec.run(args.steps)
    dt = time.perf_counter() - t0
    vec.close()
    value = n_envs * args.steps / dt
    sample = "%d of %d envs, %d steps each after a %d-step fill, %d pthreads" % (
        n_envs, ENVS_PER_GPU * args.gpus, args.steps, FILL_STEPS, threads)
    print(json.dumps({
        "impl": "reference", "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8/f64", "data": "synthetic",
        "config": {"workload": WORKLOAD % ENVS_PER_GPU, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU (default: BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from optical_rl_gym_b200 import OpticalVecEnv, nsfnet, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    n = args.envs
    env = OpticalVecEnv(ENV_ID, n, nsfnet(), device=dev, env_id_base=rank * n, seed=1, collect_info=False, **ENV_ARGS)
    actions = torch.empty((n, 1), dtype=torch.int32, device=dev)
    launches_per_step = 2                               # random_action_kernel + step_kernel

    def one_step():
        env.sample_actions(out=actions)
        env.step_raw(actions)

    for _ in range(FILL_STEPS):                         # bring the network to steady state (setup, untimed)
        one_step()
    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    # ---- timed region: K steps, CUDA events on the launching stream
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    with sampler:
        ev0.record()
        for _ in range(args.steps):
            one_step()
        ev1.record()
        torch.cuda.synchronize()
        elapsed_ms = ev0.elapsed_time(ev1)
        # keep the GPU under the same load a little longer so that the clock sampler sees it
        t_end = time.perf_counter() + 0.25
        while time.perf_counter() < t_end:
            for _ in range(200):
                one_step()
            torch.cuda.synchronize()
    stats = sharding.global_statistics(env.reduce_counters())      # the per-rollout NCCL all-reduce (8e)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = n * world * args.steps / (elapsed_ms * 1e-3)

    # ---- end to end through the public API with HOST buffers (pinned): H2D actions, step, D2H obs/reward/done.
    # Every rank runs it at the same time (they share the host's PCIe / memory), max over ranks.
    e2e_result = None
    if not args.no_e2e:
        h_act = torch.zeros((n, 1), dtype=torch.int32).pin_memory()
        h_obs = torch.zeros((n, env.obs_dim), dtype=torch.float32).pin_memory()
        h_rew = torch.zeros(n, dtype=torch.float32).pin_memory()
        h_done = torch.zeros(n, dtype=torch.uint8).pin_memory()
        h_act.copy_(actions.cpu())
        k_e2e = max(20, min(args.steps, 300))

        def e2e_step():
            d_act = h_act.to(dev, non_blocking=True)
            obs, rew, done, _ = env.step(d_act)
            h_obs.copy_(obs, non_blocking=True)
            h_rew.copy_(rew, non_blocking=True)
            h_done.copy_(done, non_blocking=True)
            torch.cuda.synchronize()            # the host policy needs the observation before it can act

        for _ in range(5):
            e2e_step()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            e2e_step()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_result = {"value": n * world * k_e2e / float(dt.item()), "unit": "env-steps/s",
                      "h2d_bytes_per_step": h_act.numel() * 4 * world,
                      "d2h_bytes_per_step": (h_obs.numel() * 4 + h_rew.numel() * 4 + h_done.numel()) * world,
                      "steps": k_e2e, "n_gpus": world,
                      "note": "VecEnv.step with pinned host buffers on every rank concurrently, synchronised every step"}

    out = None
    if rank == 0:
        # ---- roofline of the dominant kernel (step_kernel): events around each step launch only
        reps = min(args.steps, 500)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in evs:
            env.sample_actions(out=actions)
            a.record()
            env.step_raw(actions)
            b.record()
        torch.cuda.synchronize()
        step_ms = sum(a.elapsed_time(b) for a, b in evs) / reps
        peak, peak_src = read_peaks()
        achieved = B_STEP * n / (step_ms * 1e-3) / 1e9
        # in the timed region the kernels are chained by programmatic dependent launch, so the step kernel's prologue
        # runs under the previous kernel's tail; ms_per_step (action kernel included) bounds its in-stream cost
        in_stream = B_STEP * n / (elapsed_ms / args.steps * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "kernel": "deeprmsa_fast_kernel<22,5,1,false,true>", "kernel_ms": step_ms,
                    "algorithmic_bytes_per_env_step": B_STEP, "peak_source": peak_src,
                    "note": "kernel_ms = events around isolated step-kernel launches (no overlap with neighbours)",
                    "achieved_in_stream": in_stream, "frac_in_stream": in_stream / peak}
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                roofline["traffic"] = json.load(f).get("step_kernel_dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            pass

        # ---- cold-L2 variant: flush L2 (write a 256 MB buffer) before every timed step
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        cold = []
        for _ in range(50):
            env.sample_actions(out=actions)
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            env.step_raw(actions)
            b.record()
            torch.cuda.synchronize()
            cold.append(a.elapsed_time(b))
        del flush
        cold_ms = sorted(cold)[len(cold) // 2]

        e2e = e2e_result

        cpu = None if args.no_cpu_baseline else cpu_baseline(os.cpu_count() or 1)
        out = {
            "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": WORKLOAD % n, "envs_per_gpu": n, "fill_steps": FILL_STEPS, "obs": "float32 [N,54] on device",
                       "l2": "no flush: state %.0f MB/GPU (> 126 MB L2) is re-touched every step as in a rollout; "
                             "cold-L2 step time in cold_l2_ms_per_step" % (env.state_bytes / 1e6),
                       "parallelism": "env-sharded x%d, no data-path collective" % world},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "clocks": sampler.summary(), "cold_l2_ms_per_step": cold_ms,
            "accept_rate": 1.0 - stats["service_blocking_rate"], "envs_with_errors": stats["envs_with_errors"],
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
