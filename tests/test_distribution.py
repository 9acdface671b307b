"""Philox mode checked DISTRIBUTIONALLY against the reference.

The expectation is not a set of constants: tests/golden/ref_occupancy_5x4_s{000,020}.npz hold the EXACT distribution
of uniform-random lock-step play -- occupancy over the 761 observations at lock-steps t = 1 ... 256 and the per-step
probabilities of a goal for A, a goal for B and a truncation -- derived by oracle/make_occupancy.py from the
reference's own Pmat / Rmat / isd (tests/golden/ref_table_5x4_*_multi.npz, dumped from the unmodified reference).

  * chi-square of the empirical histogram of the returned observations at 13 lock-steps against `occ`
  * z-tests of the per-step goal / truncation counts against the binomial(N, p_t) they must follow
  * the statistics vector (episodes, goals, truncations, mean episode length) against its exact expectation
for slip_prob 0 and 0.2 (the registration default), the CPU oracle (-m "not gpu") and the K2 kernels (-m gpu).
"""
import os

import numpy as np
import pytest

from .conftest import GOLDEN

Z = 5.5        # two-sided z bound per check: P(|z| > 5.5) = 4e-8; a few thousand checks per run


def _fixture(slip, scenario="uniform"):
    """scenario 'uniform': both players act uniformly at random; 'policy': player A follows the fixture's table policy, B acts uniformly
    (under uniform play slip_prob does not change the distribution -- a uniform action stays uniform after slipping --
    so the 'policy' scenario is the one that pins the slip probabilities)."""
    g = np.load(os.path.join(GOLDEN, "ref_occupancy_5x4_s%03d.npz" % int(round(slip * 100))))
    if scenario == "uniform":
        return {k: g[k] for k in g.files if not k.startswith("pol_")}
    return {k[4:]: g[k] for k in g.files if k.startswith("pol_")}


def chi_square_ok(hist, p, n):
    """Pearson chi-square of a histogram of n samples against probabilities p; bins with expectation < 8 are pooled.
    Returns (statistic, dof, bound) with bound = dof + Z * sqrt(2 dof) (normal approximation of chi-square)."""
    e = p * n
    big = e >= 8
    o_b, e_b = hist[big].astype(np.float64), e[big]
    o_s, e_s = float(hist[~big].sum()), float(e[~big].sum())
    stat = float(((o_b - e_b) ** 2 / e_b).sum())
    dof = int(big.sum()) - 1
    if e_s >= 8:
        stat += (o_s - e_s) ** 2 / e_s
        dof += 1
    else:
        assert o_s <= e_s + 8 + Z * 3, "mass in bins the reference gives (almost) no probability"
    return stat, dof, dof + Z * np.sqrt(2.0 * dof)


def check_streams(fx, obs, reward, flags, stats, n, K):
    """obs / reward / flags: [K, n] arrays of a rollout that started right after reset(); stats: its statistics vector."""
    occ_t = [int(t) for t in fx["occ_t"] if t <= K]
    for row, t in enumerate(occ_t):
        hist = np.bincount(obs[t - 1].astype(np.int64), minlength=761)
        assert hist.sum() == n
        assert hist[fx["occ"][row] == 0].sum() == 0, t                 # an observation the reference cannot produce
        stat, dof, bound = chi_square_ok(hist, fx["occ"][row], n)
        assert stat < bound, (t, stat, dof, bound)
    # per-step episode ends: counts ~ binomial(n, p_t)
    term = (flags & 1) != 0
    ga = (term & (reward > 0)).sum(axis=1)
    gb = (term & (reward < 0)).sum(axis=1)
    tr = ((flags & 3) == 2).sum(axis=1)
    for name, cnt, p in (("goal_a", ga, fx["p_goal_a"]), ("goal_b", gb, fx["p_goal_b"]), ("trunc", tr, fx["p_trunc"])):
        pt = p[1:K + 1]
        sd = np.sqrt(n * pt * (1 - pt))
        z = np.where(sd > 0, (cnt - n * pt) / np.maximum(sd, 1e-300), 0.0)
        assert np.all(cnt[pt == 0] == 0), name
        assert np.abs(z).max() < Z, (name, int(np.abs(z).argmax()) + 1, float(np.abs(z).max()))
    # the statistics vector against its exact expectation; the spread is measured per env from the streams
    ep_env = ((flags & 3) != 0).sum(axis=0).astype(np.float64)
    exp_ep = n * (fx["p_goal_a"] + fx["p_goal_b"] + fx["p_trunc"])[1:K + 1].sum()
    assert stats[0] == int(ep_env.sum()) and stats[1] == int(ga.sum()) and stats[2] == int(gb.sum()) and stats[3] == int(tr.sum())
    assert stats[4] == n * K
    assert abs(stats[0] - exp_ep) <= Z * ep_env.std() * np.sqrt(n)
    exp_len = n * fx["exp_len"][1:K + 1].sum()
    # episode lengths are bounded by 100: |sum - E| < Z * 100 * sqrt(#episodes) is a (loose) sub-gaussian bound
    assert abs(stats[5] - exp_len) < Z * 100 * np.sqrt(stats[0])
    mean_len_exact = fx["exp_len"][1:K + 1].sum() / (fx["p_goal_a"] + fx["p_goal_b"] + fx["p_trunc"])[1:K + 1].sum()
    assert abs(stats[5] / stats[0] - mean_len_exact) < 0.25


def test_fixture_is_a_distribution():
    for slip in (0.0, 0.2):
        fx = _fixture(slip)
        assert np.allclose(fx["occ"].sum(axis=1), 1.0, atol=1e-9) and fx["occ"].min() >= 0
        assert int(fx["T"]) == 256 and fx["p_trunc"][:100].max() == 0 and fx["p_trunc"][100] > 0
        # symmetric game under uniform play: A and B score equally often (exactly, up to round-off)
        assert np.allclose(fx["p_goal_a"], fx["p_goal_b"], atol=1e-12)
        # lock-step 1 from the four start states: nobody can have scored
        assert fx["occ"][0][0] == 0 and fx["p_goal_a"][1] == 0
        fp = _fixture(slip, "policy")
        assert np.allclose(fp["occ"].sum(axis=1), 1.0, atol=1e-9) and fp["policy_a"].shape == (761,)
    # uniform play does not see slip_prob; table-policy play does
    assert np.allclose(_fixture(0.0)["occ"], _fixture(0.2)["occ"], atol=1e-12)
    assert np.abs(_fixture(0.0, "policy")["occ"] - _fixture(0.2, "policy")["occ"]).max() > 0.01


@pytest.mark.parametrize("scenario", ["uniform", "policy"])
@pytest.mark.parametrize("slip", [0.0, 0.2])
def test_oracle_philox_play_follows_the_reference_distribution(oracle, slip, scenario):
    """The oracle's Philox rollout (contract v2 words, decode, auto-reset) is distributed like the reference's
    play -- this pins the randomness CONTRACT itself, on the CPU."""
    n, K, seed = 1 << 16, 256, 11
    fx = _fixture(slip, scenario)
    m = oracle.OracleModel(5, 4, slip)
    init = np.array([oracle.philox_decode(oracle.philox_word(seed, i, (1 << 64) - 1))[3] for i in range(n)])
    states = np.zeros(n, oracle.STATE_DTYPE)
    for k in range(4):
        states[init == k] = m.isd[k][1]
    obs, rew, flg, st = m.rollout_philox(states, np.zeros(n, np.int32), K, seed, n_threads=8,
                                         policy_a=fx.get("policy_a"), policy_b=fx.get("policy_b"))
    check_streams(fx, obs, rew, flg, st, n, K)
    if scenario == "policy":                                   # the test has power: the other slip_prob's fixture fails
        with pytest.raises(AssertionError):
            check_streams(_fixture(0.2 - slip, scenario), obs, rew, flg, st, n, K)


@pytest.mark.gpu
@pytest.mark.parametrize("scenario", ["uniform", "policy"])
@pytest.mark.parametrize("slip,kernel", [(0.0, "table"), (0.0, "rules"), (0.2, "table"), (0.2, "rules")])
def test_k2_philox_play_follows_the_reference_distribution(slip, kernel, scenario):
    torch = pytest.importorskip("torch")
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    dev = torch.device("cuda", 0)
    n = 1 << 20 if not (kernel == "rules" and slip) else 1 << 18
    K = 256
    fx = _fixture(slip, scenario)
    env = SoccerVecEnv(n, slip_prob=slip, device=dev, rng_mode="philox", seed=20261018, kernel=kernel)
    env.reset()
    obs, rew, flg, st = env.rollout(K, policy_a=fx.get("policy_a"), policy_b=fx.get("policy_b"))
    check_streams(fx, obs.cpu().numpy(), rew.cpu().numpy(), flg.cpu().numpy(), st.cpu().numpy(), n, K)
