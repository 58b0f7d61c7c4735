"""Row f2: the reference's gym wrappers -- SimpleMatrixObservation, PathOnlyFirstFitAction,
UseInfoReward -- against golden vectors recorded from the live reference (tests/golden/make_golden_wrappers.py).
CPU: the oracle's numpy restatement; GPU (-m gpu): the device kernels through the C ABI / VecEnv wrappers.
Everything here is integer / 0-1 work: bit-exact."""
import numpy as np
import pytest

import helpers
from oracle import oracle


def _info_reward(ic, key):
    col = {"service_blocking_rate": (0, 1), "episode_service_blocking_rate": (2, 3),
           "bit_rate_blocking_rate": (4, 5), "episode_bit_rate_blocking_rate": (6, 7)}[key]
    return (ic[col[0]] - ic[col[1]]) / ic[col[0]]


@pytest.mark.parametrize("name", helpers.wrapper_golden_names())
def test_oracle_wrappers_match_reference(name):
    g = helpers.load_golden(name)
    meta = g["meta"]
    tables = helpers.golden_tables()
    for i in range(meta["n_envs"]):
        e = oracle.OracleEnv(meta["kind"], tables, **helpers.sim_kwargs(meta))
        e.set_trace(g["req_arrival"][i], g["req_holding"][i], g["req_src"][i], g["req_dst"][i], g["req_bit_rate"][i])
        e.reset(full=True)
        e.reset(full=False)
        for t in range(meta["T"]):
            if meta["matrix"]:
                obs = oracle.simple_matrix_observation(e, tables.num_nodes)
                assert obs.dtype == np.float64 and len(obs) == int(g["matrix_dim"])
                assert np.array_equal(np.packbits(obs.astype(np.uint8), bitorder="little"), g["matrix_bits"][i, t]), ("matrix", t)
            if meta["path_only"]:
                a = oracle.path_only_first_fit(e, tables, int(g["path_action"][i, t]))
                assert tuple(a) == tuple(g["actions"][i, t]), ("mapped action", t, a, g["actions"][i, t])
            else:
                a = g["actions"][i, t]
            o, rc = e.step(a)
            assert rc == 0 and o.accepted == g["accepted"][i, t], ("accepted", t)
            want = g["reward"][i, t]
            got = _info_reward([int(x) for x in o.info_counters], meta["info_key"]) if meta["info_key"] else o.reward
            assert got == want, ("reward", t, got, want)
            assert o.done == g["done"][i, t]
            if o.done:
                e.reset(full=False)
        if meta["matrix"]:
            obs = oracle.simple_matrix_observation(e, tables.num_nodes)
            assert np.array_equal(np.packbits(obs.astype(np.uint8), bitorder="little"), g["matrix_bits"][i, meta["T"]])


@pytest.mark.gpu
@pytest.mark.parametrize("name", helpers.wrapper_golden_names())
def test_cuda_wrappers_match_reference(name):
    import torch

    from optical_rl_gym_b200 import OpticalVecEnv
    from optical_rl_gym_b200.wrappers import PathOnlyFirstFitAction, SimpleMatrixObservation, UseInfoReward

    g = helpers.load_golden(name)
    meta = g["meta"]
    n, T = meta["n_envs"], meta["T"]
    base = OpticalVecEnv(meta["kind"], n, helpers.golden_tables(), traffic="trace", record_decisions=True,
                         **meta["env_args"])
    base.set_trace(g["req_arrival"], g["req_holding"], g["req_src"], g["req_dst"], g["req_bit_rate"])
    env = base
    if meta["matrix"]:
        env = SimpleMatrixObservation(env)
        assert env.observation_space.shape == (int(g["matrix_dim"]),)
    if meta["path_only"]:
        env = PathOnlyFirstFitAction(env)
        assert env.action_space.n == base.k_paths + base.reject_action
    if meta["info_key"]:
        env = UseInfoReward(env, meta["info_key"])
    base.reset(full=True)
    obs = env.reset()

    def check_matrix(o, t):
        if meta["matrix"]:
            assert o.dtype == torch.uint8
            got = np.packbits(o.cpu().numpy(), axis=1, bitorder="little")
            assert np.array_equal(got, g["matrix_bits"][:, t]), ("matrix", t)

    check_matrix(obs, 0)
    for t in range(T):
        a = g["path_action"][:, t] if meta["path_only"] else g["actions"][:, t]
        obs, reward, done, info = env.step(torch.as_tensor(a, device="cuda"))
        if meta["path_only"]:
            mapped = env._mapped if not meta["info_key"] else env.venv._mapped
            assert np.array_equal(mapped.cpu().numpy(), g["actions"][:, t]), ("mapped action", t)
        assert np.array_equal(base.decisions.cpu().numpy()[:, 0], g["accepted"][:, t]), ("accepted", t)
        assert np.array_equal(reward.cpu().numpy().astype(np.float64), g["reward"][:, t]), ("reward", t)
        assert np.array_equal(done.cpu().numpy(), g["done"][:, t]), ("done", t)
        check_matrix(obs, t + 1)
    if meta["matrix"]:      # float64 flavour = the values the reference returns
        o64 = SimpleMatrixObservation(base, dtype=torch.float64).observation()
        assert o64.dtype == torch.float64 and torch.equal(o64.to(torch.uint8), obs)
    assert int(base.error_flags().abs().sum()) == 0
    base.close()


@pytest.mark.gpu
def test_cuda_wrappers_on_wide_layout_match_oracle():
    """PathOnlyFirstFitAction + SimpleMatrixObservation beyond 32 links / 128 slots (multi-word masks, CSR hop lists)."""
    import torch

    from optical_rl_gym_b200 import OpticalVecEnv
    from optical_rl_gym_b200.topology import synthetic_ring_chords
    from optical_rl_gym_b200.wrappers import PathOnlyFirstFitAction, SimpleMatrixObservation

    tables = synthetic_ring_chords(num_nodes=20, num_chords=22, k_paths=6, seed=11)
    args = dict(episode_length=40, load=300, mean_service_holding_time=25, num_spectrum_resources=200, allow_rejection=True)
    n, T, seed = 16, 150, 9
    base = OpticalVecEnv("RMSA-v0", n, tables, traffic="philox", seed=seed, record_decisions=True, **args)
    env = PathOnlyFirstFitAction(SimpleMatrixObservation(base))
    okw = helpers.sim_kwargs(dict(kind="RMSA-v0", env_args=args))
    orc = []
    for i in range(n):
        o = oracle.OracleEnv("RMSA-v0", tables, **okw)
        o.set_philox(seed, i)
        o.reset(full=True)
        orc.append(o)
    rng = np.random.default_rng(0)
    obs = env.reset()
    n_acc = 0
    for t in range(T):
        want_obs = np.stack([oracle.simple_matrix_observation(o, tables.num_nodes) for o in orc])
        assert np.array_equal(obs.cpu().numpy(), want_obs.astype(np.uint8)), ("matrix", t)
        pa = rng.integers(0, tables.k_paths + 1, n)
        want = np.array([oracle.path_only_first_fit(o, tables, int(a)) for o, a in zip(orc, pa)])
        obs, reward, done, info = env.step(torch.as_tensor(pa, device="cuda"))
        assert np.array_equal(env._mapped.cpu().numpy(), want), ("mapped", t)
        for i, o in enumerate(orc):
            so, _ = o.step(want[i])
            assert so.accepted == int(base.decisions[i, 0]), ("accepted", t, i)
            n_acc += so.accepted
            if so.done:
                o.reset(full=False)
    assert n_acc > 0.2 * n * T
    base.close()
