"""Drive the UNMODIFIED reference (TEST INFRASTRUCTURE; build container only).

/root/reference does not exist on the GPU box, so nothing that runs there may import
this module.  It is used by oracle/make_golden.py (to write tests/golden/*.npz) and by
the `-m "not gpu"` tests that are skipped when /root/reference is absent.

How the reference is driven without touching it:
  * oracle/ref_shim provides the three `gym` names it imports (see ref_shim/gym/__init__.py);
  * `env.np_random` is replaced by ReplayRandom, whose `.random()` returns the next
    injected draw -- the reference makes exactly one `.random()` call per step()
    (SIM:395) and per reset() (SIM:414);
  * states are injected with `env.state = tuple`, the pattern the reference's own
    tests use (tests/test_deterministic_soccer_simultaneous_env.py:43).
"""
from __future__ import annotations

import os
import sys

import numpy as np

REFERENCE_ROOT = os.environ.get("SOCCER_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shim")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "gym_soccer", "envs", "soccer_simultaneous_env.py"))


def import_reference():
    """Returns the reference's SoccerSimultaneousEnv class (unmodified source)."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for p in (REFERENCE_ROOT, _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    from gym_soccer.envs.soccer_simultaneous_env import SoccerSimultaneousEnv  # noqa: E402
    return SoccerSimultaneousEnv


class ReplayRandom:
    """Stands in for np.random.RandomState: `.random()` replays injected uniform draws."""

    def __init__(self):
        self.queue = []
        self.n_calls = 0

    def push(self, u: float):
        self.queue.append(float(u))

    def random(self):
        self.n_calls += 1
        return self.queue.pop(0)

    def seed(self, seed=None):  # reset(seed=...) calls this; draws stay injected
        pass


def u_from_2bit(r) -> float:
    """Injected 2-bit value -> uniform draw: outcome r of 4, r>>1 of 2, 0 of 1 (exact in fp64)."""
    return (int(r) + 0.5) / 4.0


def u_from_u32(r) -> float:
    return (int(r) + 0.5) / 4294967296.0


def dump_model(env) -> dict:
    """Constructor products (SIM:35-144) as arrays."""
    nS = env.nS
    tuples = np.full((nS, 5), -1, np.int16)
    for st, idx in env.state_space.items():
        tuples[idx] = st
    goal = np.array([list(k) + [v] for k, v in env.goal_states.items()], dtype=np.int16)
    return dict(
        nS=np.int64(nS), nA=np.int64(env.nA), width=np.int64(env.width), height=np.int64(env.height),
        goal_rows=np.array(env.goal_rows, np.int16), goal_cols=np.array(env.goal_cols, np.int16),
        n_unreachable=np.int64(len(env.unreachable_states)), goal_states=goal,
        tuples=tuples,
        isd_prob=np.array([p for p, _ in env.isd], np.float64),
        isd_state=np.array([s for _, s in env.isd], np.int16),
        isd_obs=np.array([env._state_to_observation(s) for _, s in env.isd], np.int32),
    )


def dump_table(env) -> dict:
    """env.P and env.P_readable (SIM:258-279) as padded arrays, list order preserved."""
    nS = env.nS
    keys = sorted(env.P[1].keys())  # (aa, ab) tuples, or ints in single-agent mode
    nk = len(keys)
    assert nk in (5, 25)
    L = max(len(env.P[s][k]) for s in range(nS) for k in env.P[s].keys())
    count = np.zeros((nS, nk), np.uint8)
    prob = np.zeros((nS, nk, L), np.float64)
    nxt = np.zeros((nS, nk, L), np.int32)
    rew = np.zeros((nS, nk, L), np.int8)
    done = np.zeros((nS, nk, L), np.uint8)
    ntup = np.full((nS, nk, L, 5), -1, np.int8)
    # P[0] holds whatever key set the last-enumerated goal state left there; in single-agent
    # mode that is only the folded key (SIM:187-188 index the policy with s == 0).
    key0 = np.zeros(nk, np.uint8)
    for s in range(nS):
        for ki, k in enumerate(keys):
            if k not in env.P[s]:
                assert s == 0
                continue
            if s == 0:
                key0[ki] = 1
            tl = env.P[s][k]
            count[s, ki] = len(tl)
            for j, (p, ns, r, d) in enumerate(tl):
                prob[s, ki, j] = p
                nxt[s, ki, j] = ns
                rew[s, ki, j] = int(r)
                assert float(r) in (-1.0, 0.0, 1.0)
                done[s, ki, j] = bool(d)
    # next tuples from P_readable (goal tuples are kept as tuples there)
    names = env.ACTION_STRING
    for st, idx in env.state_space.items():
        if idx == 0:
            continue
        for ki, k in enumerate(keys):
            kk = (names[k[0]], names[k[1]]) if nk == 25 else names[k]
            for j, (p, ns, r, d) in enumerate(env.P_readable[st][kk]):
                ntup[idx, ki, j] = ns
    return dict(count=count, prob=prob, next_obs=nxt, reward=rew, done=done, next_tuple=ntup, p0_keys=key0)


def dump_dense(env) -> dict:
    """Pmat in COO form + dense Rmat (SIM:170-171, 258-279)."""
    nz = np.nonzero(env.Pmat)
    return dict(
        pmat_shape=np.array(env.Pmat.shape, np.int64),
        pmat_idx=np.stack(nz, axis=1).astype(np.int16),
        pmat_val=env.Pmat[nz].astype(np.float64),
        rmat=env.Rmat.astype(np.float64),
    )


def replay_rollout(env, act_a, act_b, rng8, rng32=None, init_rng=None):
    """Lock-step auto-reset contract, played env by env through the reference.

    Per env i: reset with draw u(init_rng[i]); then for t in range(T): step with the
    injected step draw, and -- iff terminated or truncated -- reset with the injected
    reset draw.  act_* / rng* are [T, N]; act_b is None in single-agent mode.
    Returns obs, reward (the first return agent's), flags, reset_obs, info_p, init_obs.
    """
    T, N = act_a.shape
    rr = ReplayRandom()
    env.np_random = rr
    agents = list(env.return_agent)
    a0 = agents[0]
    obs = np.zeros((T, N), np.int32)
    rew = np.zeros((T, N), np.float32)
    flg = np.zeros((T, N), np.uint8)
    rob = np.zeros((T, N), np.int32)
    inf = np.zeros((T, N), np.float64)
    init_obs = np.zeros(N, np.int32)
    for i in range(N):
        rr.push(u_from_2bit(init_rng[i] & 3))
        o, _ = env.reset()
        init_obs[i] = o[a0]
        for t in range(T):
            rr.push(u_from_u32(rng32[t, i]) if rng32 is not None else u_from_2bit(rng8[t, i] & 3))
            if env.multiagent:
                action = {"player_a": int(act_a[t, i]), "player_b": int(act_b[t, i])}
            else:
                action = {a0: int(act_a[t, i])}
            o, r, d, tr, info = env.step(action)
            if env.multiagent:
                assert o["player_a"] == o["player_b"] and r["player_b"] == -r["player_a"]
            obs[t, i] = o[a0]
            rew[t, i] = r[a0]
            flg[t, i] = (1 if d[a0] else 0) | (2 if tr[a0] else 0)
            inf[t, i] = info[a0]["p"]
            ro = o[a0]
            if d[a0] or tr[a0]:
                rr.push(u_from_2bit((rng8[t, i] >> 2) & 3))
                o2, _ = env.reset()
                ro = o2[a0]
            rob[t, i] = ro
    assert not rr.queue
    return obs, rew, flg, rob, inf, init_obs
