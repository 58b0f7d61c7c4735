"""The synthetic (counter-based Philox) traffic against what the LIVE reference draws with CPython's MT19937
(rmsa_env.py:545-561, optical_network_env.py:156-173).  Bit-for-bit parity with the reference is by trace replay; THIS
test pins the distributions: inter-arrival / holding times (two-sample Kolmogorov-Smirnov), source / destination / bit
rate (chi-square homogeneity), and the blocking rate against load under the reference's SAP-FF heuristic.  The reference
sample is tests/golden/traffic_reference_sample.npz (make_golden_traffic.py); the Philox side is the CPU oracle, which
the GPU tests hold bit-identical to the CUDA kernels (requests, actions, decisions)."""
import os

import numpy as np
import pytest
from scipy import stats

import helpers
from oracle import oracle

REF = np.load(os.path.join(helpers.GOLDEN_DIR, "traffic_reference_sample.npz"))
ALPHA = 1e-3           # per-test significance (the Philox sample is fixed by the seed: the test is deterministic)


def philox_requests(n, seed, env_index, node_prob=None):
    o = oracle.OracleEnv("DeepRMSA-v0", helpers.golden_tables(), num_slots=100, episode_length=10 ** 9, node_prob=node_prob)
    o.set_philox(seed, env_index)
    o.reset(full=True)
    out = {k: np.zeros(n, np.float64 if k in ("iat", "holding") else np.int64) for k in ("iat", "holding", "src", "dst", "bit_rate")}
    last = 0.0
    for i in range(n):
        r = o.request()
        out["iat"][i], last = r["arrival"] - last, r["arrival"]
        out["holding"][i], out["src"][i], out["dst"][i], out["bit_rate"][i] = r["holding"], r["src"], r["dst"], r["bit_rate"]
        o.step(0)
    return out


def homogeneity_p(a, b, n_cat):
    table = np.stack([np.bincount(a, minlength=n_cat), np.bincount(b, minlength=n_cat)])
    table = table[:, table.sum(0) > 0]
    return stats.chi2_contingency(table)[1]


@pytest.mark.parametrize("prefix,probs", [("", None), ("nu_", "nu_probs")])
def test_philox_requests_follow_the_reference_distributions(prefix, probs):
    node_prob = None if probs is None else REF[probs]
    n = len(REF[prefix + "iat"])
    got = philox_requests(n, seed=2024, env_index=7, node_prob=node_prob)
    for key, mean in (("iat", 0.1), ("holding", 25.0)):
        assert stats.ks_2samp(got[key], REF[prefix + key]).pvalue > ALPHA, key
        assert stats.kstest(got[key], "expon", args=(0, mean)).pvalue > ALPHA, key            # expovariate(1 / mean)
        assert abs(got[key].mean() - mean) < 4 * mean / np.sqrt(n), key
    N = 14
    assert homogeneity_p(got["src"], REF[prefix + "src"].astype(np.int64), N) > ALPHA
    assert homogeneity_p(got["dst"], REF[prefix + "dst"].astype(np.int64), N) > ALPHA
    assert homogeneity_p(got["src"] * N + got["dst"], (REF[prefix + "src"] * N + REF[prefix + "dst"]).astype(np.int64), N * N) > ALPHA
    assert homogeneity_p(got["bit_rate"], REF[prefix + "bit_rate"].astype(np.int64), 101) > ALPHA
    # exact support: randint(25, 100), src != dst, and the source marginal itself
    assert got["bit_rate"].min() == 25 and got["bit_rate"].max() == 100 and (got["src"] != got["dst"]).all()
    p = np.full(N, 1.0 / N) if node_prob is None else node_prob
    assert stats.chisquare(np.bincount(got["src"], minlength=N), p * n).pvalue > ALPHA


def test_blocking_rate_against_load_matches_the_reference():
    tables = helpers.golden_tables()
    for load, steps, acc in zip(REF["load_erlang"], REF["load_steps"], REF["load_accepted"]):
        warm, T, n_envs = int(4 * load), 2500, 48
        vec = oracle.OracleVec("DeepRMSA-v0", tables, n_envs, seed=99, num_slots=100, episode_length=10 ** 9,
                               mean_holding=25.0, mean_iat=25.0 / float(load), stats=False)
        vec.run(warm, policy=11, with_obs=False)
        got = vec.run(T, policy=11, with_obs=False) / float(n_envs * T)         # accepted requests of the T steps
        want = float(acc) / float(steps)
        vec.close()
        # the reference figure comes from 18 000 autocorrelated steps: ~0.01 standard error at 250 / 600 Erlang
        tol = 0.005 if load < 100 else 0.03
        assert abs(got - want) < tol, (load, got, want)
