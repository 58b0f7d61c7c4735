#!/usr/bin/env python
"""Statistical fixture for the synthetic (Philox) traffic: samples of what the LIVE reference draws with CPython's
MT19937 (rmsa_env.py:545-561, optical_network_env.py:156-173), plus its blocking rates at three loads.

Run in the build container only:   python tests/golden/make_golden_traffic.py   ->  traffic_reference_sample.npz
  iat / holding / src / dst / bit_rate   [R] consecutive requests of DeepRMSA-v0 envs (defaults; uniform and the
                                          non-uniform node probabilities of the reference's tests/test_deeprmsa.py)
  load_erlang [3], load_steps [3], load_accepted [3]   DeepRMSA-v0 + the reference's SAP-FF heuristic
                                          (deeprmsa_env.py:146-155), steps counted after a warm-up of 4 mean holding times
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

NONUNIFORM = [0.01801802, 0.04004004, 0.05305305, 0.01901902, 0.04504505, 0.02402402, 0.06706707,
              0.08908909, 0.13813814, 0.12212212, 0.07607608, 0.12012012, 0.01901902, 0.16916917]


def sample_requests(n_req, seed, probs):
    topo = rh.load_topology()
    env = rh.make("DeepRMSA-v0", topology=topo, seed=seed, episode_length=10 ** 9, node_request_probabilities=probs)
    out = {k: [] for k in ("iat", "holding", "src", "dst", "bit_rate")}
    last = 0.0
    for _ in range(n_req):
        s = env.current_service
        out["iat"].append(s.arrival_time - last)
        last = s.arrival_time
        out["holding"].append(s.holding_time)
        out["src"].append(s.source_id); out["dst"].append(s.destination_id); out["bit_rate"].append(s.bit_rate)
        env.step(env.action_space.n - 1 if env.allow_rejection else 0)     # any action: the request stream does not depend on it
    return out


def blocking(load, steps, warm, seed):
    from optical_rl_gym.envs.deeprmsa_env import shortest_available_path_first_fit

    topo = rh.load_topology()
    env = rh.make("DeepRMSA-v0", topology=topo, seed=seed, episode_length=10 ** 9, mean_service_holding_time=25.0,
                  mean_service_inter_arrival_time=25.0 / load)
    acc = 0
    for t in range(warm + steps):
        _, reward, _, _ = env.step(shortest_available_path_first_fit(env))
        if t >= warm:
            acc += reward > 0
    return acc


def main():
    a = sample_requests(12000, 3, None)
    b = sample_requests(12000, 4, np.array(NONUNIFORM) / np.sum(NONUNIFORM))
    out = {k: np.asarray(v) for k, v in a.items()}
    out.update({"nu_" + k: np.asarray(v) for k, v in b.items()})
    out["nu_probs"] = np.array(NONUNIFORM) / np.sum(NONUNIFORM)
    loads, steps, accs = [], [], []
    for load in (50.0, 250.0, 600.0):
        warm, n_steps, tot = int(4 * load), 6000, 0
        for seed in (11, 12, 13):
            tot += blocking(load, n_steps, warm, seed)
        loads.append(load); steps.append(3 * n_steps); accs.append(tot)
        print("load %.0f: accept %.4f over %d steps" % (load, tot / (3 * n_steps), 3 * n_steps))
    out["load_erlang"], out["load_steps"], out["load_accepted"] = np.array(loads), np.array(steps), np.array(accs)
    path = os.path.join(HERE, "traffic_reference_sample.npz")
    np.savez_compressed(path, **out)
    print(path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
