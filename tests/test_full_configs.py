"""BASELINE.json configs[3] and configs[4] at their real shapes / sizes:
  C3  RMSA-v0 on the synthetic 100-node / 300-link topology, 320 slots, k = 10 (49 500 paths), device SAP-FF
  C4  RMCSA-v0 on NSFNET, 7 cores x 320 slots, 262 144 envs on one GPU, device first-core first-fit heuristic
Every env: conservation invariants; sampled envs: bit-for-bit against the CPU oracle on the same Philox streams."""
import os

import numpy as np
import pytest

import helpers

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _c3_tables():
    from optical_rl_gym_b200.topology import TopologyTables

    return TopologyTables.load(os.path.join(helpers.GOLDEN_DIR, "topo_c3_ring100_chords200_k10.npz"))


def _compare_with_oracle(env, kind, tables, env_args, seed, T, hid, sample, with_alloc):
    from oracle import oracle

    okw = helpers.sim_kwargs(dict(kind=kind, env_args=env_args))
    m, alloc, now, nheap = env.export_state(allocation=with_alloc)
    cnt = env.counters().cpu().numpy()
    req, sid = env.current_requests()
    S = env.num_spectrum_resources
    for i in sample:
        o = oracle.OracleEnv(kind, tables, **okw)
        o.set_philox(seed, i)
        o.reset(full=True)
        o.rollout(T, policy=10 + hid)
        oa, oal, onow, onh = o.state()
        bits = (m[i].unsqueeze(-1) >> torch.arange(32, device=m.device, dtype=torch.int32)) & 1
        got = bits.reshape(m.shape[1], -1)[:, :S].to(torch.uint8).cpu().numpy()
        assert np.array_equal(got.reshape(oa.shape), oa), ("masks", i)
        if with_alloc:
            assert np.array_equal(alloc[i].cpu().numpy(), oal), ("allocation", i)
        assert now[i].item() == onow and nheap[i].item() == onh, ("clock / live services", i)
        assert np.array_equal(cnt[i], o.counters()), ("counters", i)
        r = o.request()
        assert (req["arrival"][i], req["holding"][i], req["src"][i], req["dst"][i], req["bit_rate"][i], sid[i]) == \
               (r["arrival"], r["holding"], r["src"], r["dst"], r["bit_rate"], r["service_id"]), ("pending request", i)
        o.close()


def test_config3_rmsa_100_nodes_300_links_320_slots_k10():
    from optical_rl_gym_b200 import OpticalVecEnv

    tables = _c3_tables()
    assert (tables.num_nodes, tables.num_links, tables.k_paths, tables.num_paths) == (100, 300, 10, 49500)
    n, T, seed = 4096, 1000, 5
    env_args = dict(episode_length=400, load=600, mean_service_holding_time=25, num_spectrum_resources=320, allow_rejection=True)
    env = OpticalVecEnv("RMSA-v0", n, tables, traffic="philox", seed=seed, **env_args)
    acc = torch.zeros(n, dtype=torch.int64, device="cuda")
    ndone = torch.zeros(n, dtype=torch.int64, device="cuda")
    a = torch.empty((n, 2), dtype=torch.int32, device="cuda")
    for t in range(T):
        env.heuristic("sap_ff", out=a)
        _, reward, done, _ = env.step(a)
        acc += (reward > 0)
        ndone += done
    assert int(env.error_flags().abs().sum()) == 0
    cnt = env.counters()
    assert torch.all(cnt[:, 0] == T + 1) and torch.all(cnt[:, 1] == acc)
    assert torch.all(ndone == T // 399)
    # every busy slot belongs to exactly one live service
    m, alloc, now, nheap = env.export_state(allocation=True)
    avail = env.available_slots()
    assert torch.equal((avail == 0), (alloc.reshape(avail.shape) >= 0))
    assert torch.all(nheap <= acc) and int(nheap.max()) > 300          # the network has filled up
    del m, alloc, avail
    rng = np.random.default_rng(3)
    sample = sorted(set([0, 1, 31, 32, 127, 128, n - 1] + rng.integers(0, n, 64).tolist()))
    _compare_with_oracle(env, "RMSA-v0", tables, env_args, seed, T, helpers.HEURISTIC_ID["sap_ff"], sample, with_alloc=True)
    env.close()


def test_config4_rmcsa_7_cores_320_slots_262144_envs():
    from optical_rl_gym_b200 import OpticalVecEnv

    tables = helpers.golden_tables()
    n, T, seed = 262144, 600, 8
    env_args = dict(episode_length=250, load=700, mean_service_holding_time=25, num_spectrum_resources=320,
                    num_spatial_resources=7, worst_xt=-84.7, allow_rejection=True)
    env = OpticalVecEnv("RMCSA-v0", n, tables, traffic="philox", seed=seed, **env_args)
    # a second, small handle holding the same global env ids [65536, 65536 + 2048): results must not depend on the batch
    base, ns = 65536, 2048
    small = OpticalVecEnv("RMCSA-v0", ns, tables, traffic="philox", seed=seed, env_id_base=base, **env_args)
    acc = torch.zeros(n, dtype=torch.int64, device="cuda")
    a = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    a2 = torch.empty((ns, 4), dtype=torch.int32, device="cuda")
    for t in range(T):
        env.heuristic("sap_ff", out=a)
        small.heuristic("sap_ff", out=a2)
        assert torch.equal(a[base:base + ns], a2), t
        _, reward, done, _ = env.step(a)
        _, r2, d2, _ = small.step(a2)
        assert torch.equal(reward[base:base + ns], r2) and torch.equal(done[base:base + ns], d2), t
        acc += (reward > 0)
    assert int(env.error_flags().abs().sum()) == 0 and int(small.error_flags().abs().sum()) == 0
    cnt = env.counters()
    assert torch.all(cnt[:, 0] == T) and torch.all(cnt[:, 1] == acc)        # RMCSA counts in step (rmcsa_env.py:292-293)
    assert torch.all(cnt[:, 4] >= 2 * cnt[:, 5])                            # bit_rate_requested is counted twice (App. B-7)
    mb, _, nowb, nhb = env.export_state()
    ms, alloc_s, nows, nhs = small.export_state(allocation=True)
    assert torch.equal(mb[base:base + ns], ms) and torch.equal(nowb[base:base + ns], nows) and torch.equal(nhb[base:base + ns], nhs)
    assert torch.equal(cnt[base:base + ns], small.counters())
    avail_s = small.available_slots()
    assert torch.equal((avail_s == 0), (alloc_s.reshape(avail_s.shape) >= 0))   # busy slots == slots of live services
    # busy-slot total of every env == what its counters imply is possible: 0 < busy <= cores * links * slots
    busy = (mb.to(torch.int64) & 0xFFFFFFFF)
    assert int(nhb.min()) >= 0 and int(nhb.max()) <= env.heap_capacity
    del busy
    rate = float(acc.sum()) / (n * T)
    assert 0.2 < rate < 0.8, rate            # the reach test (_crosstalk_is_acceptable) blocks about half of the requests
    rng = np.random.default_rng(4)
    sample = sorted(set([0, 31, 32, n - 1] + rng.integers(0, n, 64).tolist()))
    _compare_with_oracle(env, "RMCSA-v0", tables, env_args, seed, T, helpers.HEURISTIC_ID["heuristic"], sample, with_alloc=False)
    env.close(); small.close()
