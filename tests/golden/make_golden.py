#!/usr/bin/env python
"""Records golden vectors from the LIVE, unmodified reference (/root/reference).

Run in the build container only (the GPU box has no reference tree):
    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

For every case it drives real reference envs (via refshim + ref_harness) with a seeded
action source and dumps, per env and step t (request t is the one action t answers):
  req_arrival/req_holding/req_src/req_dst/req_bit_rate/req_id   [n_env, T+1]
  actions [n_env, T, A]; accepted, path_row, initial_slot, number_slots, core, mod [n_env, T]
  reward [n_env, T] f64; done [n_env, T] u8; counters [n_env, T, 8] i64 (after the step)
  info_* [n_env, T] f64 (every float key of the reference's info dict)
  obs [n_env, T+1, obs_dim] f64 (DeepRMSA; obs[0] = after construction)
  avail_bits [n_env, T, C*E, ceil(S/8)] u8 (np.packbits, little bit order) for env 0..n_mask_envs-1
  final_avail [n_env, C, E, S] i8, final_alloc [n_env, C, E, S] i32
plus the env_args needed to rebuild the run.  The episode driver is
utils.evaluate_heuristic's (reset() at the start of every episode).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "..", "optical-rl-gym_b200"))

import ref_harness as rh  # noqa: E402

from optical_rl_gym_b200.topology import TopologyTables  # noqa: E402

NONUNIFORM = [0.01801802, 0.04004004, 0.05305305, 0.01901902, 0.04504505, 0.02402402, 0.06706707,
              0.08908909, 0.13813814, 0.12212212, 0.07607608, 0.12012012, 0.01901902, 0.16916917]


def path_row_of(tables, env, path):
    """Global row of a reference Path object in our tables (path_id is the row)."""
    return int(path.path_id)


def counters_of(env, kind):
    br = kind != "RWA-v0"
    return [env.services_processed, env.services_accepted, env.episode_services_processed,
            env.episode_services_accepted,
            env.bit_rate_requested if br else 0, env.bit_rate_provisioned if br else 0,
            env.episode_bit_rate_requested if br else 0, env.episode_bit_rate_provisioned if br else 0]


def avail_of(env, kind):
    key = "available_wavelengths" if kind == "RWA-v0" else "available_slots"
    a = np.asarray(env.topology.graph[key])
    return a.reshape((-1,) + a.shape[-2:]) if a.ndim == 3 else a[None]


def alloc_of(env, kind):
    a = env.spectrum_wavelengths_allocation if kind == "RWA-v0" else env.spectrum_slots_allocation
    a = np.asarray(a)
    return a if a.ndim == 3 else a[None]


def make_policy(kind, name, env, rng):
    """Seeded action sources.  'mixed' exercises accepts, busy-slot rejections, out-of-range
    indices and the explicit reject action; the named heuristics are the reference's own."""
    k, S = env.k_paths, env.num_spectrum_resources
    if kind == "DeepRMSA-v0":
        from optical_rl_gym.envs import deeprmsa_env as m
        n_act = k * env.j + env.reject_action
        return {"random": lambda: int(rng.integers(0, n_act)),
                "random_oob": lambda: int(rng.integers(0, k * env.j + 2)),
                "sp": lambda: m.shortest_path_first_fit(env),
                "sap": lambda: m.shortest_available_path_first_fit(env)}[name]
    if kind == "RMSA-v0":
        from optical_rl_gym.envs import rmsa_env as m

        def mixed():
            u = rng.random()
            if u < 0.45:       # first fit (incl. the last feasible start) on a random path
                p = int(rng.integers(0, k))
                path = env.k_shortest_paths[env.current_service.source, env.current_service.destination][p]
                n = env.get_number_slots(path)
                for s in range(0, S - n + 1):
                    if env.is_path_free(path, s, n):
                        return (p, s)
                return (p, int(rng.integers(0, S)))
            if u < 0.9:        # uniformly random incl. out-of-range path / slot
                return (int(rng.integers(0, k + 1)), int(rng.integers(0, S + 1)))
            return (k, S)
        return {"mixed": mixed, "sp_ff": lambda: m.shortest_path_first_fit(env),
                "sap_ff": lambda: m.shortest_available_path_first_fit(env),
                "llp_ff": lambda: m.least_loaded_path_first_fit(env)}[name]
    if kind == "RWA-v0":
        from optical_rl_gym.envs import rwa_env as m

        def mixed():
            u = rng.random()
            if u < 0.6:
                p = int(rng.integers(0, k))
                path = env.k_shortest_paths[env.current_service.source, env.current_service.destination][p]
                for w in range(S):
                    if env.is_path_free(path, w):
                        return (p, w)
                return (p, int(rng.integers(0, S)))
            if u < 0.95:
                return (int(rng.integers(0, k + 1)), int(rng.integers(0, S + 1)))
            return (k, S)
        return {"mixed": mixed, "sp_ff": lambda: m.shortest_path_first_fit(env),
                "sap_ff": lambda: m.shortest_available_path_first_fit(env),
                "sap_lf": lambda: m.shortest_available_path_last_fit(env),
                "llp_ff": lambda: m.least_loaded_path_first_fit(env)}[name]
    if kind == "RMCSA-v0":
        from optical_rl_gym.envs import rmcsa_env as m
        M, Cc = len(env.modulation_formats), env.num_spatial_resources

        def heuristic():
            a = m.shortest_available_path_best_modulation_first_core_first_fit(env)
            return a if len(a) == 4 else (k, M, Cc, S)   # the reference's reject tuple has 3 entries (would raise)

        def mixed():
            u = rng.random()
            if u < 0.5:
                p, mod, core = int(rng.integers(0, k)), int(rng.integers(0, min(M, 3))), int(rng.integers(0, Cc))
                path = env.k_shortest_paths[env.current_service.source, env.current_service.destination][p]
                n = env.get_number_slots(path, env.modulation_formats[mod])
                for s in range(0, S - n + 1):
                    if env.is_path_free(path, core, s, n):
                        return (p, mod, core, s)
                return (p, mod, core, int(rng.integers(0, S)))
            if u < 0.9:
                return (int(rng.integers(0, k + 1)), int(rng.integers(0, M)), int(rng.integers(0, Cc + 1)),
                        int(rng.integers(0, S + 1)))
            return (k, M - 1, Cc, S)
        return {"mixed": mixed, "heuristic": heuristic}[name]
    raise KeyError(kind)


def record(kind, env_args, policy, n_envs, T, seed0, n_mask_envs=1):
    topo = rh.load_topology()
    tables = TopologyTables.from_graph(topo)
    adim = {"RWA-v0": 2, "RMSA-v0": 2, "DeepRMSA-v0": 1, "RMCSA-v0": 4}[kind]
    out = {}

    def put(name, i, t, value, shape, dtype):
        if name not in out:
            out[name] = np.zeros(shape, dtype)
        out[name][i, t] = value

    for i in range(n_envs):
        args = dict(env_args)
        if "node_request_probabilities" in args and args["node_request_probabilities"] is not None:
            args["node_request_probabilities"] = np.array(args["node_request_probabilities"])
        env = rh.make(kind, topology=topo, seed=seed0 + i, **args)
        rng = np.random.default_rng(1000 + seed0 + i)
        pol = make_policy(kind, policy, env, rng)
        S = env.num_spectrum_resources
        Cc = getattr(env, "num_spatial_resources", 1)
        E = env.topology.number_of_edges()

        def put_req(t):
            s = env.current_service
            put("req_arrival", i, t, s.arrival_time, (n_envs, T + 1), np.float64)
            put("req_holding", i, t, s.holding_time, (n_envs, T + 1), np.float64)
            put("req_src", i, t, s.source_id, (n_envs, T + 1), np.int32)
            put("req_dst", i, t, s.destination_id, (n_envs, T + 1), np.int32)
            put("req_bit_rate", i, t, s.bit_rate if s.bit_rate is not None else 0, (n_envs, T + 1), np.int32)
            put("req_id", i, t, s.service_id, (n_envs, T + 1), np.int32)

        obs = env.reset()
        stats_slot = {}
        put_req(0)
        if kind == "DeepRMSA-v0":
            put("obs", i, 0, obs, (n_envs, T + 1, len(obs)), np.float64)
        for t in range(T):
            a = pol()
            svc = env.current_service
            obs, reward, done, info = env.step(a)
            put("actions", i, t, np.atleast_1d(np.asarray(a, dtype=np.int64))[:adim], (n_envs, T, adim), np.int32)
            put("accepted", i, t, int(svc.accepted), (n_envs, T), np.uint8)
            acc = bool(svc.accepted)
            put("path_row", i, t, path_row_of(tables, env, svc.path) if acc else -1, (n_envs, T), np.int32)
            if kind == "RWA-v0":
                put("initial_slot", i, t, svc.wavelength if acc else -1, (n_envs, T), np.int32)
                put("number_slots", i, t, 1 if acc else -1, (n_envs, T), np.int32)
            else:
                put("initial_slot", i, t, svc.initial_slot if acc else -1, (n_envs, T), np.int32)
                put("number_slots", i, t, svc.number_slots if acc else -1, (n_envs, T), np.int32)
            if kind == "RMCSA-v0":
                put("core", i, t, svc.core if acc else -1, (n_envs, T), np.int32)
                put("mod", i, t, env.modulation_formats.index(svc.current_modulation) if acc else -1,
                    (n_envs, T), np.int32)
            put("reward", i, t, reward, (n_envs, T), np.float64)
            put("done", i, t, int(done), (n_envs, T), np.uint8)
            put("counters", i, t, counters_of(env, kind), (n_envs, T, 8), np.int64)
            for key, val in info.items():
                if np.ndim(val) == 0:
                    put("info_" + key, i, t, float(val), (n_envs, T), np.float64)
                elif key in ("path_action_probability", "wavelength_action_probability"):      # rwa_env.py:148-151
                    val = np.asarray(val, np.float64)
                    put("info_" + key, i, t, val, (n_envs, T, len(val)), np.float64)
            if i < n_mask_envs and (t < 48 or t % 16 == 15 or t == T - 1):
                # row f1: the time-averaged statistics the reference keeps ON the topology graph (never returned by step):
                # per link utilization / external_fragmentation / compactness (rmsa_env.py:464-543, rmcsa_env.py:591-688,
                # rwa_env.py:365-383) and the graph's throughput / compactness (rmsa_env.py:439-462, rmcsa_env.py:560-589),
                # by link index; snapshots after steps 0..47, then every 16th (running averages: any error persists)
                ls = np.zeros((E, 3), np.float64)
                for n1, n2 in env.topology.edges():
                    d = env.topology[n1][n2]
                    ls[d["index"]] = (d.get("utilization", 0.0), d.get("external_fragmentation", 0.0), d.get("compactness", 0.0))
                ts = stats_slot.setdefault(t, len(stats_slot))
                n_snap = 48 + len([x for x in range(48, T) if x % 16 == 15 or x == T - 1])
                put("graph_link_stats", i, ts, ls, (n_mask_envs, n_snap, E, 3), np.float64)
                put("graph_stats", i, ts, (env.topology.graph.get("throughput", 0.0), env.topology.graph.get("compactness", 0.0)),
                    (n_mask_envs, n_snap, 2), np.float64)
                put("graph_stats_step", 0, ts, t, (1, n_snap), np.int32)
            if i < n_mask_envs:
                bits = np.packbits(avail_of(env, kind).reshape(Cc * E, S).astype(np.uint8), axis=1, bitorder="little")
                put("avail_bits", i, t, bits, (n_mask_envs, T, Cc * E, bits.shape[1]), np.uint8)
            if done:
                obs = env.reset()
            put_req(t + 1)
            if kind == "DeepRMSA-v0":
                put("obs", i, t + 1, obs, (n_envs, T + 1, len(obs)), np.float64)
        if "final_avail" not in out:
            out["final_avail"] = np.zeros((n_envs, Cc, E, S), np.int8)
            out["final_alloc"] = np.zeros((n_envs, Cc, E, S), np.int32)
            out["final_now"] = np.zeros(n_envs, np.float64)
            out["final_nheap"] = np.zeros(n_envs, np.int32)
        out["final_avail"][i] = avail_of(env, kind)
        out["final_alloc"][i] = alloc_of(env, kind)
        out["final_now"][i] = env.current_time
        out["final_nheap"][i] = len(env._events)
    meta = dict(kind=kind, env_args=env_args, policy=policy, n_envs=n_envs, T=T, seed0=seed0)
    out["meta"] = np.array(json.dumps(meta))
    return out


CASES = {
    # name: (kind, env_args, policy, n_envs, T, seed0)
    "deeprmsa_default_random": ("DeepRMSA-v0", dict(episode_length=50), "random", 4, 1500, 10),
    "deeprmsa_default_sap": ("DeepRMSA-v0", dict(episode_length=50), "sap", 2, 1500, 20),
    "deeprmsa_j3_nonuniform_rej": ("DeepRMSA-v0", dict(
        episode_length=40, j=3, allow_rejection=True, mean_service_holding_time=7.5,
        mean_service_inter_arrival_time=1.0 / 12.0, node_request_probabilities=NONUNIFORM), "random_oob", 3, 1200, 30),
    "deeprmsa_j2_sp_s64": ("DeepRMSA-v0", dict(episode_length=30, j=2, num_spectrum_resources=64), "sp", 2, 800, 40),
    "rmsa_load250_mixed": ("RMSA-v0", dict(episode_length=50, load=250, mean_service_holding_time=25,
                                           allow_rejection=True), "mixed", 4, 1500, 50),
    "rmsa_load250_sap_ff": ("RMSA-v0", dict(episode_length=50, load=250, mean_service_holding_time=25,
                                            allow_rejection=True), "sap_ff", 2, 1500, 60),
    "rmsa_load600_llp_ff": ("RMSA-v0", dict(episode_length=100, load=600, mean_service_holding_time=25,
                                            allow_rejection=True), "llp_ff", 2, 1000, 70),
    "rmsa_s64_discrete_sp_ff": ("RMSA-v0", dict(episode_length=100, load=50, mean_service_holding_time=25,
                                                num_spectrum_resources=64, bit_rate_selection="discrete",
                                                allow_rejection=True), "sp_ff", 2, 1000, 80),
    "rwa_load450_mixed": ("RWA-v0", dict(episode_length=100, load=450, mean_service_holding_time=25), "mixed", 3, 2000, 90),
    "rwa_load450_sap_ff": ("RWA-v0", dict(episode_length=1000, load=450, mean_service_holding_time=25), "sap_ff", 2, 2500, 10),
    "rwa_load450_sap_lf": ("RWA-v0", dict(episode_length=200, load=450, mean_service_holding_time=25), "sap_lf", 1, 1500, 11),
    "rwa_load450_llp_ff": ("RWA-v0", dict(episode_length=200, load=450, mean_service_holding_time=25), "llp_ff", 1, 1500, 12),
    "rwa_load300_sp_ff": ("RWA-v0", dict(episode_length=200, load=300, mean_service_holding_time=25), "sp_ff", 1, 1500, 13),
    "rmcsa_s64_heuristic": ("RMCSA-v0", dict(episode_length=100, load=250, mean_service_holding_time=25,
                                             num_spectrum_resources=64, num_spatial_resources=7, worst_xt=-84.7,
                                             allow_rejection=True), "heuristic", 2, 1500, 10),
    "rmcsa_s64_mixed": ("RMCSA-v0", dict(episode_length=100, load=700, mean_service_holding_time=25,
                                         num_spectrum_resources=64, num_spatial_resources=7, worst_xt=-84.7,
                                         allow_rejection=True), "mixed", 2, 1500, 20),
    "rmcsa_s100_c3_mixed": ("RMCSA-v0", dict(episode_length=60, load=400, mean_service_holding_time=25,
                                             num_spectrum_resources=100, num_spatial_resources=3, worst_xt=-84.7,
                                             allow_rejection=True), "mixed", 2, 1000, 30),
}


def main(argv):
    names = argv[1:] or list(CASES)
    topo = rh.load_topology()
    TopologyTables.from_graph(topo).save(os.path.join(HERE, "nsfnet_tables.npz"))
    for name in names:
        kind, env_args, policy, n_envs, T, seed0 = CASES[name]
        out = record(kind, env_args, policy, n_envs, T, seed0)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print("%-32s %8.1f KB  accept=%.3f" % (name, os.path.getsize(path) / 1024, out["accepted"].mean()))


if __name__ == "__main__":
    main(sys.argv)
