#!/usr/bin/env python
"""Device-side throughput of the OTHER BASELINE.json configs (parity-test cases, not bench lines):
C1 RMSA-v0 NSFNET 4096 envs SAP-FF; C3 RMSA-v0 100-node/300-link synthetic graph, 320 slots, k=10, SAP-FF;
C4 RMCSA-v0 NSFNET 7 cores x 320 slots, first-core first-fit; plus RWA-v0 (C0's env) batched.
    python tools/bench_configs.py [c1 c3 c4 rwa]        -> one JSON line per config
Every step = heuristic kernel (device action source) + step kernel; CUDA events, state resident in HBM.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "optical-rl-gym_b200"))
import torch  # noqa: E402

from optical_rl_gym_b200 import OpticalVecEnv, nsfnet, synthetic_ring_chords  # noqa: E402

CONFIGS = {
    "c1": ("RMSA-v0", "nsfnet", 4096, "sap_ff", dict(episode_length=1000, load=250, mean_service_holding_time=25)),
    "c1x": ("RMSA-v0", "nsfnet", 65536, "sap_ff", dict(episode_length=1000, load=250, mean_service_holding_time=25)),
    "rwa": ("RWA-v0", "nsfnet", 65536, "sap_ff", dict(episode_length=1000, load=450, mean_service_holding_time=25)),
    "c3": ("RMSA-v0", "ring100", 131072, "sap_ff", dict(episode_length=1000, load=600, mean_service_holding_time=25,
                                                        num_spectrum_resources=320)),
    "c4": ("RMCSA-v0", "nsfnet", 262144, "heuristic", dict(episode_length=1000, load=800, mean_service_holding_time=25,
                                                           num_spectrum_resources=320, num_spatial_resources=7,
                                                           worst_xt=-84.7)),
}


def main():
    names = sys.argv[1:] or ["c1", "c1x", "rwa", "c3", "c4"]
    for name in names:
        kind, topo, n, heur, args = CONFIGS[name]
        t0 = time.time()
        tables = nsfnet() if topo == "nsfnet" else synthetic_ring_chords(100, 200, k_paths=10, seed=1)
        t_topo = time.time() - t0
        env = OpticalVecEnv(kind, n, tables, seed=1, collect_info=False, **args)
        hid = "sap_ff"
        a = torch.empty((n, env.action_dim), dtype=torch.int32, device="cuda")
        fill, steps = int(os.environ.get("FILL", 400)), int(os.environ.get("STEPS", 200))

        def one():
            env.heuristic(hid, out=a)
            env.step_raw(a)

        for _ in range(fill):
            one()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        cnt = env.counters().sum(0).cpu().numpy()
        print(json.dumps({"config": name, "env": kind, "topology": "%s (%d nodes, %d links, k=%d)" % (
            tables.name, tables.num_nodes, tables.num_links, tables.k_paths), "envs": n, "args": args,
            "policy": "device " + heur, "us_per_step": ms * 1e3, "env_steps_per_s": n / (ms * 1e-3),
            "accept_rate": float(cnt[1]) / float(cnt[0]), "state_MB": env.state_bytes / 1e6,
            "envs_with_errors": int((env.error_flags() != 0).sum()), "topology_seconds": round(t_topo, 1)}), flush=True)
        env.close()
        del env
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
