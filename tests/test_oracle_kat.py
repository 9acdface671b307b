"""Known-answer transitions from the reference's own test-suite, run against the CPU oracle.
Each case cites /root/reference/gym_soccer/tests/test_deterministic_soccer_simultaneous_env.py
("TD") or test_general.py ("TG").  CPU only.  When /root/reference is present the same cases are
also run against the unmodified reference so that the expectations themselves are pinned."""
import numpy as np
import pytest

NOOP, N, S, E, W = 0, 1, 2, 3, 4

# (state, aa, ab, expected next states (set, any draw), reward_a or None, terminated)
KAT = [
    ((1, 5, 3, 1, 0), E, NOOP, None, 1.0, True),                      # TD:49, TD:411-414
    ((3, 5, 1, 1, 1), NOOP, W, None, -1.0, True),                     # TD:52
    ((1, 1, 3, 5, 0), W, NOOP, None, -1.0, True),                     # TD:56-60 A own goal
    ((2, 1, 3, 5, 0), W, NOOP, None, -1.0, True),                     # TD:64-68
    ((3, 1, 1, 5, 1), NOOP, E, None, 1.0, True),                      # TD:72-76 B own goal
    ((3, 1, 2, 5, 1), NOOP, E, None, 1.0, True),                      # TD:80-84
    ((1, 2, 1, 3, 0), E, W, {(1, 2, 1, 3, 0), (1, 2, 1, 3, 1)}, 0.0, False),   # TD:89-100 swap
    ((1, 2, 1, 3, 1), E, W, {(1, 2, 1, 3, 0), (1, 2, 1, 3, 1)}, 0.0, False),
    ((1, 2, 1, 3, 0), E, NOOP, {(1, 2, 1, 3, 1)}, 0.0, False),        # TD:105-109 standing collision
    ((1, 2, 1, 3, 1), NOOP, W, {(1, 2, 1, 3, 0)}, 0.0, False),        # TD:112-116
    ((0, 1, 3, 5, 0), N, E, {(0, 1, 3, 5, 0)}, 0.0, False),           # TD:170-264 walls
    ((0, 1, 3, 5, 1), W, S, {(0, 1, 3, 5, 1)}, 0.0, False),
    ((3, 5, 0, 1, 0), S, N, {(3, 5, 0, 1, 0)}, 0.0, False),
    ((3, 5, 0, 1, 1), E, W, {(3, 5, 0, 1, 1)}, 0.0, False),
    ((1, 1, 3, 3, 1), W, NOOP, {(1, 1, 3, 3, 1)}, 0.0, False),        # TD:269-321 goal mouth, no ball
    ((2, 5, 3, 3, 1), E, NOOP, {(2, 5, 3, 3, 1)}, 0.0, False),
    ((3, 3, 1, 5, 0), NOOP, E, {(3, 3, 1, 5, 0)}, 0.0, False),
    ((3, 3, 2, 1, 0), NOOP, W, {(3, 3, 2, 1, 0)}, 0.0, False),
    ((1, 1, 3, 3, 0), E, W, {(1, 2, 3, 2, 0)}, 0.0, False),           # TD:333-339
    ((1, 1, 3, 3, 1), E, W, {(1, 2, 3, 2, 1)}, 0.0, False),
    ((1, 5, 1, 1, 0), E, W, None, 1.0, True),                         # TD:343-352 only the holder scores
    ((1, 5, 1, 1, 1), E, W, None, -1.0, True),
    ((1, 1, 1, 2, 0), E, E, {(1, 2, 1, 3, 0)}, 0.0, False),           # TD:356-371 follow into vacated cell
    ((1, 1, 1, 3, 1), E, E, {(1, 2, 1, 4, 1)}, 0.0, False),
    ((0, 1, 3, 4, 1), N, E, {(0, 1, 3, 5, 1)}, 0.0, False),           # TD:404-407
]
FOUR_WAY = [((1, 1, 2, 2, 0), E, N), ((1, 2, 2, 1, 1), W, N), ((2, 1, 1, 2, 0), E, S), ((2, 2, 1, 1, 1), W, S),
            ((1, 1, 1, 3, 0), E, W), ((1, 3, 1, 1, 1), W, E), ((1, 1, 3, 1, 0), S, N), ((3, 1, 1, 1, 1), N, S)]  # TD:146-165


def _check(step_all, is_goal):
    for st, aa, ab, want, rew, term in KAT:
        seen = set()
        for r in range(4):
            ns, reward, done = step_all(st, aa, ab, (r + 0.5) / 4)
            assert done == term and reward == rew, (st, aa, ab)
            if term:
                assert is_goal(ns)
            seen.add(ns)
        if want is not None:
            assert seen == want, (st, aa, ab, seen)
    for st, aa, ab in FOUR_WAY:
        outs = [step_all(st, aa, ab, (r + 0.5) / 4)[0] for r in range(4)]
        assert len(set(outs)) == 4
        assert sum(o[:2] != st[:2] for o in outs) == 2 and sum(o[2:4] != st[2:4] for o in outs) == 2
        assert sum(o[4] != st[4] for o in outs) == 2


def test_oracle_known_answers(oracle):
    m = oracle.OracleModel(5, 4, 0.0)
    assert m.nS == 761 and m.nA == 5 and m.width == 7            # TG:175-180, TD:16
    assert [m.state_to_obs(s) for _, s in m.isd] == [253, 254, 435, 436]

    def step_all(st, aa, ab, u):
        e = oracle.OracleEnv(m)
        e.reset(0.1)
        e.state = st
        o, r, d, t, p = e.step(aa * 5 + ab, u)
        return e.state, r, d
    _check(step_all, m.is_goal_state)


def test_oracle_reset_and_truncation(oracle):
    m = oracle.OracleModel(5, 4, 0.0)
    e = oracle.OracleEnv(m)
    with pytest.raises(AssertionError):
        e.step(0, 0.5)                                            # SIM:376
    for u, want in [(0.0, 253), (0.2499, 253), (0.25, 254), (0.5, 435), (0.75, 436), (0.999, 436)]:
        assert e.reset(u)[0] == want                              # strict > in categorical_sample
    e.state = (0, 1, 3, 5, 0)
    for t in range(100):
        o, r, d, tr, p = e.step(0, 0.5)
        assert tr == (t == 99) and not d                          # SIM:404
    with pytest.raises(AssertionError):
        e.step(0, 0.5)


def test_reference_agrees_with_the_expectations():
    from oracle import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("/root/reference is only present in the build container")
    Env = rh.import_reference()
    env = Env(5, 4, 0.0)
    rr = rh.ReplayRandom()
    env.np_random = rr

    def step_all(st, aa, ab, u):
        rr.push(0.1)
        env.reset()
        env.state = st
        rr.push(u)
        o, r, d, t, info = env.step({'player_a': aa, 'player_b': ab})
        return tuple(env.state), r['player_a'], d['player_a']
    _check(step_all, lambda s: s in env.goal_states)
    assert np.isclose(sum(p for p, _ in env.isd), 1.0)
