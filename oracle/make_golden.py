"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python oracle/make_golden.py            # all fixtures
    python oracle/make_golden.py --quick    # skip the 11x7 pitch (its constructor takes ~30 s)

Fixtures (all produced by /root/reference code, never by the oracle or the CUDA path):
  ref_table_<w>x<h>_s<slip>_<mode>.npz   constructor products, full P / P_readable dump,
                                         Pmat (COO) and Rmat
  ref_rollout_<w>x<h>_s<slip>_<mode>.npz injected-randomness auto-reset rollouts
                                         (inputs and outputs)
mode: multi | a_free (player_b_policy folded) | b_free (player_a_policy folded)

The inputs are regenerated from fixed numpy RandomState seeds and stored alongside the
outputs, so the tests never depend on numpy's stream staying stable.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def slip_tag(s):
    return f"{int(round(s * 100)):03d}"


def random_policy(nS, seed):
    """Same construction as the reference's utils/policies.py:4-9 get_random_policy."""
    rs = np.random.RandomState(seed)
    return {s: int(rs.randint(0, 5)) for s in range(nS)}


def make_env(width, height, slip, mode, pol_seed=0):
    Env = rh.import_reference()
    f = width * height
    nS = 1 + 2 * f * (f - 1)
    pol = random_policy(nS, pol_seed)
    if mode == "multi":
        return Env(width=width, height=height, slip_prob=slip), None
    if mode == "a_free":
        return Env(width=width, height=height, slip_prob=slip, player_b_policy=pol), pol
    if mode == "b_free":
        return Env(width=width, height=height, slip_prob=slip, player_a_policy=pol), pol
    raise ValueError(mode)


def gen(width, height, slip, mode, T, N, dense=True):
    t0 = time.time()
    env, pol = make_env(width, height, slip, mode)
    tag = f"{width}x{height}_s{slip_tag(slip)}_{mode}"
    d = {}
    d.update(rh.dump_model(env))
    d.update(rh.dump_table(env))
    if dense:
        d.update(rh.dump_dense(env))
    d["slip_prob"] = np.float64(slip)
    if pol is not None:
        d["policy"] = np.array([pol[s] for s in range(env.nS)], np.int8)
    np.savez_compressed(os.path.join(OUT, f"ref_table_{tag}.npz"), **d)
    t1 = time.time()

    rs = np.random.RandomState(1000 + width * 100 + height * 10 + int(slip * 100))
    act_a = rs.randint(0, 5, (T, N)).astype(np.uint8)
    act_b = rs.randint(0, 5, (T, N)).astype(np.uint8) if mode == "multi" else None
    rng8 = rs.randint(0, 16, (T, N)).astype(np.uint8)
    rng32 = rs.randint(0, 2 ** 32, (T, N), dtype=np.uint64).astype(np.uint32) if slip > 0 else None
    init_rng = rs.randint(0, 4, N).astype(np.uint8)
    obs, rew, flg, rob, inf, init_obs = rh.replay_rollout(env, act_a, act_b, rng8, rng32, init_rng)
    r = dict(act_a=act_a, rng8=rng8, init_rng=init_rng, obs=obs, reward=rew, flags=flg,
             reset_obs=rob, info_p=inf, init_obs=init_obs, slip_prob=np.float64(slip))
    if act_b is not None:
        r["act_b"] = act_b
    if rng32 is not None:
        r["rng32"] = rng32
    if pol is not None:
        r["policy"] = d["policy"]
    np.savez_compressed(os.path.join(OUT, f"ref_rollout_{tag}.npz"), **r)
    print(f"{tag}: nS={env.nS} table {t1 - t0:.1f}s rollout {time.time() - t1:.1f}s "
          f"(episodes={int((flg != 0).sum())})", flush=True)


def gen_native(width, height, slip, mode, seed, T):
    """The reference driven through its OWN np.random.RandomState(seed): the trajectory a user
    of the reference sees.  The drop-in single-env wrapper must reproduce it draw for draw."""
    Env = rh.import_reference()
    f = width * height
    pol = random_policy(1 + 2 * f * (f - 1), 0)
    kw = {}
    if mode == "a_free":
        kw["player_b_policy"] = pol
    elif mode == "b_free":
        kw["player_a_policy"] = pol
    env = Env(width=width, height=height, slip_prob=slip, seed=seed, **kw)
    a0 = env.return_agent[0]
    rs = np.random.RandomState(4242 + seed)
    acts = rs.randint(0, 5, (T, 2)).astype(np.uint8)
    obs = np.zeros(T, np.int32); rew = np.zeros(T, np.float32); flg = np.zeros(T, np.uint8)
    inf = np.zeros(T, np.float64); tup = np.zeros((T, 5), np.int16); rob = np.zeros(T, np.int32)
    o, _ = env.reset()
    init_obs = o[a0]
    for t in range(T):
        action = {"player_a": int(acts[t, 0]), "player_b": int(acts[t, 1])} if env.multiagent \
            else {a0: int(acts[t, 0])}
        o, r, d, tr, info = env.step(action)
        obs[t], rew[t], inf[t] = o[a0], r[a0], info[a0]["p"]
        flg[t] = (1 if d[a0] else 0) | (2 if tr[a0] else 0)
        tup[t] = env.state
        rob[t] = o[a0]
        if d[a0] or tr[a0]:
            o2, _ = env.reset()
            rob[t] = o2[a0]
    tag = f"{width}x{height}_s{slip_tag(slip)}_{mode}"
    d = dict(seed=np.int64(seed), acts=acts, obs=obs, reward=rew, flags=flg, info_p=inf, state=tup,
             reset_obs=rob, init_obs=np.int32(init_obs), slip_prob=np.float64(slip))
    if mode != "multi":
        d["policy"] = np.array([pol[s] for s in range(env.nS)], np.int8)
    np.savez_compressed(os.path.join(OUT, f"ref_native_{tag}.npz"), **d)
    print(f"native {tag}: episodes={int((flg != 0).sum())}", flush=True)


def gen_planner(slip, mode, theta=1e-10, gamma=0.99):
    """The reference's planners (utils/planners.py) on a single-agent env: value iteration over
    env.P and modified policy iteration over env.Pmat / env.Rmat."""
    sys.path.insert(0, rh.REFERENCE_ROOT)
    env, pol = make_env(5, 4, slip, mode)
    from gym_soccer.utils.planners import (modified_policy_iteration, policy_evaluation, policy_improvement,
                                           policy_iteration, value_iteration)
    t0 = time.time()
    pi, V, Q, cc = value_iteration(env, theta=theta, discount_factor=gamma)
    t1 = time.time()
    mpi, mV, mQ, mcc = modified_policy_iteration(env, k=1, theta=theta, discount_factor=gamma)
    t2 = time.time()
    # the list-walking planners on a fixed (seeded) policy: evaluation, one improvement, full policy iteration
    pe_pi = np.random.RandomState(99).randint(0, env.nA, env.nS)
    pe_V = policy_evaluation(pe_pi, env, theta=1e-8, discount_factor=gamma)
    imp_pi, imp_Q = policy_improvement(pe_V, env, discount_factor=gamma)
    np.random.seed(4321)
    it_pi, it_V, it_Q, it_cc = policy_iteration(env, theta=1e-8, discount_factor=gamma)
    tag = f"5x4_s{slip_tag(slip)}_{mode}"
    np.savez_compressed(os.path.join(OUT, f"ref_planner_{tag}.npz"), theta=theta, gamma=gamma,
                        policy=np.array([pol[s] for s in range(env.nS)], np.int8),
                        vi_pi=pi, vi_V=V, vi_Q=Q, vi_cc=cc, mpi_pi=mpi, mpi_V=mV, mpi_Q=mQ, mpi_cc=mcc,
                        pe_pi=pe_pi, pe_V=pe_V, pe_theta=1e-8, imp_pi=imp_pi, imp_Q=imp_Q,
                        it_seed=4321, it_pi=it_pi, it_V=it_V, it_Q=it_Q, it_cc=it_cc,
                        vi_seconds=t1 - t0, mpi_seconds=t2 - t1, pi_seconds=time.time() - t2)
    print(f"planner {tag}: VI {cc} sweeps {t1 - t0:.1f}s, MPI {mcc} iterations {t2 - t1:.1f}s, "
          f"PI {it_cc} iterations {time.time() - t2:.1f}s", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--native-only", action="store_true")
    ap.add_argument("--planner-only", action="store_true")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    if args.planner_only:
        gen_planner(0.2, "a_free")
        gen_planner(0.0, "b_free")
        return
    gen_native(5, 4, 0.0, "multi", seed=0, T=3000)
    gen_native(5, 4, 0.2, "multi", seed=7, T=3000)
    gen_native(5, 4, 0.2, "a_free", seed=3, T=1500)
    gen_native(5, 4, 0.2, "b_free", seed=5, T=1500)
    gen_native(7, 5, 0.2, "multi", seed=11, T=1500)
    if args.native_only:
        return
    gen(5, 4, 0.0, "multi", T=1500, N=64)
    gen(5, 4, 0.2, "multi", T=1500, N=64)
    gen(5, 4, 0.2, "a_free", T=600, N=32)
    gen(5, 4, 0.2, "b_free", T=600, N=32)
    gen(5, 4, 0.0, "a_free", T=600, N=32)
    gen(6, 4, 0.0, "multi", T=600, N=32)
    gen(7, 5, 0.0, "multi", T=600, N=32)
    gen(7, 5, 0.2, "multi", T=400, N=16, dense=False)
    if not args.quick:
        gen(9, 6, 0.0, "multi", T=300, N=16, dense=False)
        gen(11, 7, 0.0, "multi", T=300, N=16, dense=False)


if __name__ == "__main__":
    main()
