#!/usr/bin/env python
"""Golden vectors for SURVEY.md row f2 (wrappers), recorded from the LIVE, unmodified reference.

Build container only (needs /root/reference):   python tests/golden/make_golden_wrappers.py
Writes tests/golden/wrap_*.npz.  Per case, for every env i and step t:
  req_* [n, T+1]              the request the action of step t answers (trace replay input)
  path_action [n, T]          what the agent gave the PathOnlyFirstFitAction wrapper
  actions [n, T, 2]           what the wrapper turned it into (rmsa_env.py:840-874, rwa_env.py:505-536)
  accepted [n, T], reward [n, T] f64, done [n, T]
  matrix_bits [n, T+1, ceil(D/8)]  SimpleMatrixObservation output (rmsa_env.py:806-837, rmcsa_env.py:914-947),
                              a 0/1 float64 vector of length D, packed with np.packbits(bitorder="little")
  info_reward [n, T]          UseInfoReward(env, key) reward (wrappers.py:4-16)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "..", "optical-rl-gym_b200"))

import ref_harness as rh  # noqa: E402


def record(kind, env_args, n_envs, T, seed0, matrix, path_only, info_key=None):
    topo = rh.load_topology()
    if kind == "RMSA-v0":
        from optical_rl_gym.envs import rmsa_env as m
    elif kind == "RWA-v0":
        from optical_rl_gym.envs import rwa_env as m
    else:
        from optical_rl_gym.envs import rmcsa_env as m
    from optical_rl_gym.wrappers import UseInfoReward
    out = {}

    def put(name, i, t, value, shape, dtype):
        if name not in out:
            out[name] = np.zeros(shape, dtype)
        out[name][i, t] = value

    for i in range(n_envs):
        base = rh.make(kind, topology=topo, seed=seed0 + i, **env_args)
        env = base
        if matrix:
            env = m.SimpleMatrixObservation(env)
        if path_only:
            env = m.PathOnlyFirstFitAction(env)
        if info_key:
            env = UseInfoReward(env, info_key)
        rng = np.random.default_rng(7000 + seed0 + i)
        k, S = base.k_paths, base.num_spectrum_resources

        def put_req(t):
            s = base.current_service
            put("req_arrival", i, t, s.arrival_time, (n_envs, T + 1), np.float64)
            put("req_holding", i, t, s.holding_time, (n_envs, T + 1), np.float64)
            put("req_src", i, t, s.source_id, (n_envs, T + 1), np.int32)
            put("req_dst", i, t, s.destination_id, (n_envs, T + 1), np.int32)
            put("req_bit_rate", i, t, s.bit_rate if s.bit_rate is not None else 0, (n_envs, T + 1), np.int32)

        def put_obs(t, obs):
            if matrix:
                assert obs.dtype == np.float64 and set(np.unique(obs)) <= {0.0, 1.0}
                bits = np.packbits(obs.astype(np.uint8), bitorder="little")
                put("matrix_bits", i, t, bits, (n_envs, T + 1, len(bits)), np.uint8)
                out["matrix_dim"] = np.array(len(obs))

        obs = base.reset()
        if matrix:
            obs = _matrix_of(env)
        put_req(0)
        put_obs(0, obs)
        for t in range(T):
            svc = base.current_service
            if path_only:
                a = int(rng.integers(0, k + 1))                 # includes the reject action k
                mapped = _path_only_of(env).action(a)
                put("path_action", i, t, a, (n_envs, T), np.int32)
            else:                                                # RMCSA: the reference heuristic or a uniformly random 4-tuple
                a = m.shortest_available_path_best_modulation_first_core_first_fit(base) if rng.random() < 0.6 else ()
                if len(a) != 4:                                  # (the reference's reject tuple has 3 entries: would raise)
                    a = (int(rng.integers(0, k + 1)), int(rng.integers(0, 3)),
                         int(rng.integers(0, base.num_spatial_resources + 1)), int(rng.integers(0, S + 1)))
                mapped = a
            put("actions", i, t, np.asarray(mapped, np.int64), (n_envs, T, len(mapped)), np.int32)
            obs, reward, done, info = env.step(a)
            put("accepted", i, t, int(svc.accepted), (n_envs, T), np.uint8)
            put("reward", i, t, float(reward), (n_envs, T), np.float64)
            put("done", i, t, int(done), (n_envs, T), np.uint8)
            if done:
                base.reset()
                if matrix:
                    obs = _matrix_of(env)
            put_req(t + 1)
            put_obs(t + 1, obs)
    meta = dict(kind=kind, env_args=env_args, n_envs=n_envs, T=T, seed0=seed0, matrix=matrix, path_only=path_only,
                info_key=info_key)
    out["meta"] = np.array(json.dumps(meta))
    return out


def _find(env, cls_name):
    e = env
    while e is not None:
        if type(e).__name__ == cls_name:
            return e
        e = e.__dict__.get("env")
    raise KeyError(cls_name)


def _matrix_of(env):
    return _find(env, "SimpleMatrixObservation").observation(None)


def _path_only_of(env):
    return _find(env, "PathOnlyFirstFitAction")


CASES = {
    "wrap_rmsa_matrix_pathonly": ("RMSA-v0", dict(episode_length=50, load=250, mean_service_holding_time=25,
                                                  allow_rejection=True), 2, 400, 300, True, True, None),
    "wrap_rmsa_pathonly_inforeward": ("RMSA-v0", dict(episode_length=80, load=400, mean_service_holding_time=25,
                                                      allow_rejection=True, num_spectrum_resources=64), 2, 600, 310, False, True,
                                      "episode_bit_rate_blocking_rate"),
    "wrap_rwa_pathonly": ("RWA-v0", dict(episode_length=100, load=450, mean_service_holding_time=25), 2, 800, 320,
                          False, True, None),
    "wrap_rmcsa_matrix": ("RMCSA-v0", dict(episode_length=60, load=400, mean_service_holding_time=25,
                                           num_spectrum_resources=64, num_spatial_resources=3, worst_xt=-84.7,
                                           allow_rejection=True), 1, 300, 330, True, False, None),
}


def main(argv):
    for name in argv[1:] or list(CASES):
        out = record(CASES[name][0], *CASES[name][1:])
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print("%-32s %8.1f KB  accept rate %.3f" % (name, os.path.getsize(path) / 1024.0, out["accepted"].mean()))


if __name__ == "__main__":
    main(sys.argv)
