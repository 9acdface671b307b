"""Pin the CPU oracle (oracle/soccer_oracle.c) against golden vectors produced by the
UNMODIFIED reference (oracle/make_golden.py -> tests/golden/*.npz).  CPU only."""
import numpy as np
import pytest

from .conftest import golden_tags, load_golden, parse_tag


def _model(oracle, tag, g):
    w, h, slip, mode = parse_tag(tag)
    pol = None if mode == "multi" else {s: int(a) for s, a in enumerate(g["policy"])}
    kw = {}
    if mode == "a_free":
        kw["player_b_policy"] = pol
    elif mode == "b_free":
        kw["player_a_policy"] = pol
    return oracle.OracleModel(w, h, slip, **kw)


@pytest.mark.parametrize("tag", golden_tags("table"))
def test_constructor_products(oracle, tag):
    """SIM:35-165: nS, padded width, goal rows, state enumeration order, isd."""
    g = load_golden("table", tag)
    m = _model(oracle, tag, g)
    assert m.nS == int(g["nS"]) and m.nA == int(g["nA"])
    assert m.width == int(g["width"]) and m.height == int(g["height"])
    assert m.goal_rows == tuple(int(x) for x in g["goal_rows"])
    assert m.n_unreachable == int(g["n_unreachable"])
    assert m.n_goal_states == len(g["goal_states"])
    for row in g["goal_states"]:
        st = tuple(int(v) for v in row[:5])
        assert m.is_goal_state(st) and m.goal_reward(st) == float(row[5])
        assert m.state_to_obs(st) == 0
    tuples = g["tuples"]
    for obs in range(1, m.nS):
        st = tuple(int(v) for v in tuples[obs])
        assert m.obs_to_state(obs) == st
        assert m.state_to_obs(st) == obs
    assert m.obs_to_state(0) == (-1, -1, -1, -1, -1)
    assert len(m.isd) == len(g["isd_prob"])
    for k, (p, st) in enumerate(m.isd):
        assert p == g["isd_prob"][k] and st == tuple(int(v) for v in g["isd_state"][k])
        assert m.state_to_obs(st) == int(g["isd_obs"][k])


@pytest.mark.parametrize("tag", golden_tags("table"))
def test_transition_table_bit_exact(oracle, tag):
    """SIM:167-293: every P[s][key] list -- length, order, fp64 probabilities (==), next
    observation, next tuple (P_readable), reward, done."""
    g = load_golden("table", tag)
    m = _model(oracle, tag, g)
    d = m.dump_table()
    L = g["prob"].shape[2]
    assert d["prob"].shape[2] == L
    assert np.array_equal(d["count"], g["count"])
    assert np.array_equal(d["prob"], g["prob"])          # bit-exact fp64
    assert np.array_equal(d["next_obs"], g["next_obs"])
    assert np.array_equal(d["reward"], g["reward"])
    assert np.array_equal(d["done"], g["done"])
    # P_readable tuples exist for obs >= 1 only in the golden dump
    assert np.array_equal(d["next_tuple"][1:], g["next_tuple"][1:])


@pytest.mark.parametrize("tag", [t for t in golden_tags("table") if "pmat_val" in load_golden("table", t)])
def test_dense_pmat_rmat_bit_exact(oracle, tag):
    """SIM:170-171, 258-279 including the Pmat[0, 0] accumulation over all goal states."""
    g = load_golden("table", tag)
    m = _model(oracle, tag, g)
    P, R = m.pmat_rmat()
    assert tuple(g["pmat_shape"]) == P.shape
    idx = g["pmat_idx"].astype(np.int64)
    nz = np.nonzero(P)
    assert len(nz[0]) == len(idx)
    assert np.array_equal(np.stack(nz, axis=1), idx)
    assert np.array_equal(P[nz], g["pmat_val"])
    assert np.array_equal(R, g["rmat"])


@pytest.mark.parametrize("tag", golden_tags("rollout"))
def test_injected_rollout_bit_exact(oracle, tag):
    """SIM:375-424 under the auto-reset contract: obs, reward, terminated, truncated and
    the post-reset observation, step by step, against the replayed reference."""
    g = load_golden("rollout", tag)
    m = _model(oracle, tag, g)
    T, N = g["act_a"].shape
    init = g["init_rng"]
    states = np.zeros(N, oracle.STATE_DTYPE)
    for i in range(N):
        e = oracle.OracleEnv(m)
        o, _ = e.reset((int(init[i] & 3) + 0.5) / 4.0)
        assert o == int(g["init_obs"][i])
        states[i] = e.state
    ts = np.zeros(N, np.int32)
    act_b = g["act_b"] if "act_b" in g else None
    rng32 = np.ascontiguousarray(g["rng32"]) if "rng32" in g else None
    obs, rew, flg, rob = m.rollout_injected(states, ts, np.ascontiguousarray(g["act_a"]),
                                            None if act_b is None else np.ascontiguousarray(act_b),
                                            np.ascontiguousarray(g["rng8"]), rng32, n_threads=2)
    assert np.array_equal(obs, g["obs"])
    assert np.array_equal(rew, g["reward"])
    assert np.array_equal(flg, g["flags"])
    assert np.array_equal(rob, g["reset_obs"])
    assert (flg != 0).sum() > 20  # the fixture really exercises resets
