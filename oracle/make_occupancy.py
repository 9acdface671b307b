"""Exact distribution of uniform-random lock-step play, derived from the REFERENCE's own transition matrices.

    python oracle/make_occupancy.py        # reads tests/golden/ref_table_5x4_s{000,020}_multi.npz, writes
                                           # tests/golden/ref_occupancy_5x4_s{000,020}.npz

TEST INFRASTRUCTURE.  Input: `Pmat`, `Rmat`, `isd` exactly as the unmodified reference built them
(/root/reference/gym_soccer/envs/soccer_simultaneous_env.py:170-171, 258-279, 146-165; dumped by
oracle/make_golden.py).  Nothing here comes from the oracle port or the CUDA path.

Process (what SoccerVecEnv.rollout() does in Philox mode, and what a Python loop over the reference's step()/reset()
with np.random actions does): every env starts from reset() (isd), both players pick one of the 5 actions uniformly
every step (joint action uniform over 25), the env resets whenever terminated or truncated (timestep >= 100,
SIM:404).  The state of one env is (observation, timestep); its distribution is propagated exactly in fp64:

    occ[t][o]        P(the observation returned by lock-step t is o), o = 0 is a goal (SIM:493), for t in `occ_t`
    p_goal_a[t], p_goal_b[t], p_trunc[t]
                     P(lock-step t ends an episode with reward +1 / with reward -1 / by truncation without a goal)
    p_len[t][l]      not stored; exp_len[t] = E[length of the episode that ends at lock-step t] * P(it ends)

for t = 1 .. T; the same again, under keys prefixed `pol_`, for the scenario in which player A follows a fixed
table policy (`pol_policy_a`) and player B acts uniformly -- there slip_prob changes the distribution.  From these, the expected statistics vector of K <= T steps of N envs is N * the cumulative sums.
The GPU test chi-square-tests the empirical histograms of K2's observation stream against `occ` and z-tests the
per-step episode-end counts against the binomial these probabilities define.
"""
from __future__ import annotations

import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
T = 256
OCC_T = (1, 2, 3, 5, 10, 25, 50, 99, 100, 101, 150, 200, 256)
MAX_T = 100          # SIM:404


def table_policy(nS):
    """Player A's fixed table policy in the `pol_` scenario (utils/policies.py:4-9 construction, seed 101)."""
    return np.random.RandomState(101).randint(0, 5, nS).astype(np.int8)


def derive(tag, with_policies=False):
    g = np.load(os.path.join(GOLDEN, f"ref_table_{tag}_multi.npz"))
    nS = int(g["nS"])
    shape = tuple(int(x) for x in g["pmat_shape"])
    assert shape == (nS, nS, 5, 5)
    P = np.zeros(shape, np.float64)
    idx = g["pmat_idx"].astype(np.int64)
    P[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]] = g["pmat_val"]
    R = g["rmat"].astype(np.float64)                                  # E[reward | s, aa, ab] (player A's)
    P = P.reshape(nS, nS, 25)
    R = R.reshape(nS, 25)
    # every non-terminal row of Pmat is a probability distribution for every joint action
    assert np.allclose(P[1:].sum(axis=1), 1.0, atol=1e-12)
    pol = {}
    if with_policies:
        # player A follows a fixed table policy (SIM:187-188 semantics: the action is the policy at the CURRENT
        # observation), player B acts uniformly: A's chosen action is slipped, so slip_prob matters here -- under
        # uniform play it does not (a uniformly chosen action stays uniform after slipping)
        pa = table_policy(nS)
        P4, R4 = P.reshape(nS, nS, 5, 5), R.reshape(nS, 5, 5)
        M = P4[np.arange(nS), :, pa, :].mean(axis=2)
        net = R4[np.arange(nS), pa, :].mean(axis=1)
        pol = dict(policy_a=pa)
    else:
        M = P.mean(axis=2)                                            # uniform joint action
        net = R.mean(axis=1)                                          # P(+1 | s) - P(-1 | s)
    goal = M[:, 0]                                                    # P(goal | s)
    ga, gb = (goal + net) / 2, (goal - net) / 2
    isd = np.zeros(nS)
    isd[g["isd_obs"]] = g["isd_prob"]
    D = np.zeros((MAX_T, nS))                                         # D[tau][s]: timestep tau, observation s
    D[0] = isd
    occ, p_ga, p_gb, p_tr, e_len = {}, np.zeros(T + 1), np.zeros(T + 1), np.zeros(T + 1), np.zeros(T + 1)
    for t in range(1, T + 1):
        V = D @ M                                                     # [tau][next obs], column 0 = goal
        if t in OCC_T:
            occ[t] = V.sum(axis=0)
        p_ga[t] = float((D @ ga).sum())
        p_gb[t] = float((D @ gb).sum())
        trunc_mass = float(V[MAX_T - 1, 1:].sum())                    # timestep 99 -> 100 without a goal
        p_tr[t] = trunc_mass
        goal_by_tau = V[:, 0]
        e_len[t] = float((goal_by_tau * (np.arange(MAX_T) + 1)).sum() + trunc_mass * MAX_T)
        ended = float(goal_by_tau.sum()) + trunc_mass
        nD = np.zeros_like(D)
        nD[1:] = V[:-1]
        nD[:, 0] = 0.0                                                # goals left the field
        nD[0] = isd * ended                                           # reset() (SIM:410-424)
        D = nD
        assert abs(D.sum() - 1.0) < 1e-9
    return dict(T=T, occ_t=np.array(sorted(occ)), occ=np.stack([occ[t] for t in sorted(occ)]),
                p_goal_a=p_ga, p_goal_b=p_gb, p_trunc=p_tr, exp_len=e_len, slip_prob=float(g["slip_prob"]),
                source=f"ref_table_{tag}_multi.npz (Pmat / Rmat / isd of the unmodified reference)", **pol)


def main():
    for tag in ("5x4_s000", "5x4_s020"):
        d = derive(tag)
        d.update({"pol_" + k: v for k, v in derive(tag, with_policies=True).items()})
        out = os.path.join(GOLDEN, f"ref_occupancy_{tag}.npz")
        np.savez_compressed(out, **d)
        ep = d["p_goal_a"] + d["p_goal_b"] + d["p_trunc"]
        print(tag, "-> %s" % os.path.basename(out),
              "| long-run: %.3f steps / episode, %.2f %% truncated, A/B wins %.4f" % (
                  1.0 / ep[-50:].mean(), 100 * d["p_trunc"][-100:].sum() / ep[-100:].sum(),
                  d["p_goal_a"][-50:].sum() / d["p_goal_b"][-50:].sum()))


if __name__ == "__main__":
    main()
