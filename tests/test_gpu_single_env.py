"""The single-env drop-in (gym_soccer_littman94_b200.envs.SoccerSimultaneousEnv) exercised the
way the reference's own tests exercise the reference class:
  * known-answer transitions restated from
    /root/reference/gym_soccer/tests/test_deterministic_soccer_simultaneous_env.py (cited per test)
  * API shape / Python types from /root/reference/gym_soccer/tests/test_general.py
  * slip statistics from /root/reference/gym_soccer/tests/test_slip_soccer_simultaneous_env.py
  * and, beyond what the reference pins: the whole trajectory for a given seed equals the
    reference's (tests/golden/ref_native_*.npz), draw for draw.
"""
import numpy as np
import pytest

from .conftest import parse_tag

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

NATIVE = ["5x4_s000_multi", "5x4_s020_multi", "5x4_s020_a_free", "5x4_s020_b_free", "7x5_s020_multi"]


@pytest.fixture(scope="module")
def Env():
    assert torch.cuda.is_available()
    from gym_soccer_littman94_b200.envs import SoccerSimultaneousEnv
    return SoccerSimultaneousEnv


@pytest.fixture
def env(Env):
    e = Env(width=5, height=4, slip_prob=0.0)
    e.reset()
    return e


def _step(env, state, a, b):
    env.reset()
    env.state = state
    return env.step({'player_a': a, 'player_b': b})


# ---- reference trajectory for a seed, draw for draw
@pytest.mark.parametrize("tag", NATIVE)
def test_same_seed_same_trajectory_as_reference(Env, tag):
    g = np.load(f"{__import__('os').path.dirname(__file__)}/golden/ref_native_{tag}.npz")
    w, h, slip, mode = parse_tag(tag)
    kw = {}
    if mode != "multi":
        pol = {s: int(a) for s, a in enumerate(g["policy"])}
        kw["player_b_policy" if mode == "a_free" else "player_a_policy"] = pol
    env = Env(width=w, height=h, slip_prob=slip, seed=int(g["seed"]), **kw)
    a0 = env.return_agent[0]
    o, info = env.reset()
    assert o[a0] == int(g["init_obs"])
    acts = g["acts"]
    for t in range(len(acts)):
        action = {'player_a': int(acts[t, 0]), 'player_b': int(acts[t, 1])} if env.multiagent else {a0: int(acts[t, 0])}
        o, r, d, tr, info = env.step(action)
        assert o[a0] == g["obs"][t] and r[a0] == g["reward"][t], t
        assert (int(d[a0]) | (int(tr[a0]) << 1)) == g["flags"][t], t
        assert info[a0]["p"] == g["info_p"][t], t
        assert env.state == tuple(int(v) for v in g["state"][t]), t
        if d[a0] or tr[a0]:
            o2, _ = env.reset()
            assert o2[a0] == g["reset_obs"][t]


# ---- test_deterministic...:14-37
def test_initialization_and_api_shape(env):
    assert env.width == 7 and env.height == 4 and env.slip_prob == 0.0
    assert env.action_space['player_a'].n == 5 and env.action_space['player_b'].n == 5
    assert env.nS == 761 and env.nA == 5                     # test_general.py:175-180
    obs, info = env.reset()
    assert isinstance(obs, dict) and set(obs) == {'player_a', 'player_b'} and set(info) == {'player_a', 'player_b'}
    out = env.step({'player_a': env.NOOP, 'player_b': env.NOOP})
    assert len(out) == 5 and all(isinstance(x, dict) for x in out)
    obs, rew, term, trunc, info = out
    # test_general.py:299-301: plain Python types
    assert isinstance(obs['player_a'], int) and isinstance(rew['player_a'], float)
    assert isinstance(term['player_a'], bool) and isinstance(trunc['player_b'], bool)
    assert rew['player_b'] == 0 and np.signbit(rew['player_b'])     # reward * -1 -> -0.0 (SIM:402)


# ---- test_deterministic...:39-84
def test_scoring_and_own_goals(env):
    for st, a, b, ra in [((1, 5, 3, 1, 0), env.EAST, env.NOOP, 1.0), ((3, 5, 1, 1, 1), env.NOOP, env.WEST, -1.0),
                         ((1, 1, 3, 5, 0), env.WEST, env.NOOP, -1.0), ((2, 1, 3, 5, 0), env.WEST, env.NOOP, -1.0),
                         ((3, 1, 1, 5, 1), env.NOOP, env.EAST, 1.0), ((3, 1, 2, 5, 1), env.NOOP, env.EAST, 1.0)]:
        obs, rew, term, trunc, info = _step(env, st, a, b)
        assert term['player_a'] and term['player_b']
        assert rew['player_a'] == ra and rew['player_b'] == -ra
        assert obs['player_a'] == 0 and env.needs_reset
        assert env.state in env.goal_states and env.goal_states[env.state] == ra


# ---- test_deterministic...:86-116
def test_swap_and_standing_collisions(env):
    for p in (0, 1):
        seen = set()
        for _ in range(40):
            _step(env, (1, 2, 1, 3, p), env.EAST, env.WEST)
            assert env.state[:4] == (1, 2, 1, 3)
            seen.add(env.state[4])
        assert seen == {0, 1}
    _step(env, (1, 2, 1, 3, 0), env.EAST, env.NOOP)
    assert env.state == (1, 2, 1, 3, 1)          # possession flips to the stander (SIM:330-335)
    _step(env, (1, 2, 1, 3, 0), env.NOOP, env.WEST)
    assert env.state == (1, 2, 1, 3, 1)          # B walks into standing A, who held the ball: B takes it


# ---- test_deterministic...:118-165
@pytest.mark.parametrize("state,a,b", [((1, 1, 2, 2, 0), 3, 1), ((1, 2, 2, 1, 1), 4, 1), ((1, 1, 1, 3, 0), 3, 4),
                                       ((3, 1, 1, 1, 1), 1, 2)])
def test_same_target_cell_four_way(env, state, a, b):
    n = 600
    moved = {'A': 0, 'B': 0}
    switched = 0
    for _ in range(n):
        obs, rew, term, trunc, info = _step(env, state, a, b)
        assert info['player_a']['p'] == 0.25
        if env.state[:2] != state[:2]:
            moved['A'] += 1
        elif env.state[2:4] != state[2:4]:
            moved['B'] += 1
        switched += env.state[4] != state[4]
    assert moved['A'] + moved['B'] == n
    assert 0.42 <= moved['A'] / n <= 0.58 and 0.42 <= switched / n <= 0.58


# ---- test_deterministic...:167-321 walls and goal mouths
def test_walls_and_goal_mouth_bounces(env):
    for p in (0, 1):
        for a, b in [(env.NORTH, env.EAST), (env.NORTH, env.SOUTH), (env.WEST, env.EAST), (env.WEST, env.SOUTH)]:
            _step(env, (0, 1, 3, 5, p), a, b)
            assert env.state == (0, 1, 3, 5, p)
            _step(env, (3, 5, 0, 1, p), env.SOUTH if a == env.NORTH else env.EAST, env.NORTH if b == env.SOUTH else env.WEST)
            assert env.state == (3, 5, 0, 1, p)
    # goal mouth without the ball: stays (SIM:369-372)
    for st, a, b in [((1, 1, 3, 3, 1), env.WEST, env.NOOP), ((2, 1, 3, 3, 1), env.WEST, env.NOOP),
                     ((1, 5, 3, 3, 1), env.EAST, env.NOOP), ((3, 3, 1, 5, 0), env.NOOP, env.EAST),
                     ((3, 3, 2, 1, 0), env.NOOP, env.WEST)]:
        obs, rew, term, trunc, info = _step(env, st, a, b)
        assert env.state == st and not term['player_a']


# ---- test_deterministic...:331-371, 396-407
def test_free_moves_and_follow_into_vacated_cell(env):
    for p in (0, 1):
        _step(env, (1, 1, 3, 3, p), env.EAST, env.WEST)
        assert env.state == (1, 2, 3, 2, p)
        _step(env, (1, 1, 1, 2, p), env.EAST, env.EAST)
        assert env.state == (1, 2, 1, 3, p)
        _step(env, (1, 1, 1, 3, p), env.EAST, env.EAST)
        assert env.state == (1, 2, 1, 4, p)
    obs, rew, *_ = _step(env, (1, 5, 1, 1, 0), env.EAST, env.WEST)
    assert rew['player_a'] == 1
    obs, rew, *_ = _step(env, (1, 5, 1, 1, 1), env.EAST, env.WEST)
    assert rew['player_a'] == -1
    _step(env, (0, 1, 3, 4, 1), env.NORTH, env.EAST)
    assert env.state == (0, 1, 3, 5, 1)


def test_asserts_match_reference_messages(Env):
    env = Env()
    with pytest.raises(AssertionError, match="Please reset the environment before taking a step"):
        env.step({'player_a': 0, 'player_b': 0})
    env.reset()
    with pytest.raises(AssertionError, match="Action must be a dictionary"):
        env.step([0, 0])
    with pytest.raises(AssertionError, match="A policy for player_b must be provided"):     # SIM:381 fires first
        env.step({'player_a': 0})
    with pytest.raises(AssertionError, match="Action must be a dictionary of length 1 or 2"):
        env.step({'player_a': 0, 'player_b': 0, 'referee': 0})
    with pytest.raises(AssertionError, match="Both players cannot have a policy"):
        Env(player_a_policy={}, player_b_policy={})
    with pytest.raises(AssertionError, match="Width must be at least 5"):
        Env(width=4)
    with pytest.raises(AssertionError, match="Height must be at least 4"):
        Env(height=3)
    # truncation after 100 steps, then the assert again (SIM:404-406)
    env.reset()
    env.state = (0, 1, 3, 5, 0)
    for t in range(100):
        obs, rew, term, trunc, info = env.step({'player_a': 0, 'player_b': 0})
        assert trunc['player_a'] == (t == 99) and not term['player_a']
    with pytest.raises(AssertionError, match="Please reset"):
        env.step({'player_a': 0, 'player_b': 0})


# ---- test_general.py:93-156
def test_reset_distribution(env):
    counts = {}
    for _ in range(2000):
        env.reset()
        counts[env.state] = counts.get(env.state, 0) + 1
    assert set(counts) == {s for _, s in env.isd}
    assert np.std(list(counts.values())) / np.mean(list(counts.values())) < 0.08


# ---- test_general.py:159-261
def test_single_agent_modes(Env):
    pol = {s: int(a) for s, a in enumerate(np.random.RandomState(0).randint(0, 5, 761))}
    for free, kw in (("player_a", dict(player_b_policy=pol)), ("player_b", dict(player_a_policy=pol))):
        env = Env(width=5, height=4, slip_prob=0.2, **kw)
        other = "player_b" if free == "player_a" else "player_a"
        assert not env.multiagent and env.return_agent == [free]
        assert env.observation_space[free].n == 761 and other not in env.observation_space
        assert env.action_space[free].n == 5 and other not in env.action_space
        obs, info = env.reset()
        assert set(obs) == {free} and set(info) == {free} and 0 <= obs[free] < 761
        out = env.step({free: 3})
        assert all(set(d) == {free} for d in out)
        with pytest.raises(AssertionError):
            env.step({other: 1})


# ---- test_slip...:39-59, 198-210
def test_slip_statistics(Env):
    env = Env(width=5, height=4, slip_prob=0.2, seed=1)
    n, scored = 3000, 0
    for _ in range(n):
        env.reset()
        env.state = (1, 5, 3, 1, 0)
        obs, rew, term, trunc, info = env.step({'player_a': env.EAST, 'player_b': env.NOOP})
        scored += term['player_a']
    assert abs(scored / n - 0.8) < 0.03
    for _ in range(200):                     # NOOP never slips
        env.reset()
        env.state = (1, 2, 2, 4, 0)
        env.step({'player_a': env.NOOP, 'player_b': env.NOOP})
        assert env.state == (1, 2, 2, 4, 0)


def test_render_smoke(env, capsys):
    env.render()
    out = capsys.readouterr().out
    assert "Player A position" in out and "Ball possession" in out


def test_make_factory():
    import gym_soccer_littman94_b200 as pkg
    e = pkg.make("SoccerSimultaneous-v0")
    assert e.slip_prob == 0.2 and e.width == 7          # gym_soccer/__init__.py:5-12 (commented block)
    v = pkg.make("SoccerSimultaneous-v0", num_envs=8, slip_prob=0.0)
    assert v.num_envs == 8


# ---- the speculative step (soccer_step_speculate: every (joint action, draw) of the current state in one launch,
#      enqueued before the action is known) returns what the launch-and-wait step returns
@pytest.mark.parametrize("slip", [0.0, 0.2])
@pytest.mark.parametrize("w,h,mode", [(5, 4, "multi"), (5, 4, "a_free"), (5, 4, "b_free"), (7, 5, "multi"), (6, 4, "b_free")])
def test_speculative_step_equals_launch_and_wait(Env, monkeypatch, w, h, mode, slip):
    """slip 0: all 25 joint actions x 4 draw values per launch; slip 0.2: the 25 joint actions with the draw the env's
    generator is going to make (shadow generator one draw ahead) -- also when the caller draws from, reseeds or
    replaces np_random in between."""
    rs = np.random.RandomState(7)
    kw = {}
    probe = Env(width=w, height=h)
    if mode != "multi":
        pol = {s: int(a) for s, a in enumerate(rs.randint(0, 5, probe.nS))}
        kw["player_b_policy" if mode == "a_free" else "player_a_policy"] = pol
    envs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("SOCCER_B200_SINGLE_ENV_SPECULATE", flag)
        envs.append(Env(width=w, height=h, slip_prob=slip, seed=11, **kw))
    spec, plain = envs
    assert spec._spec_on and not plain._spec_on
    assert spec.reset() == plain.reset()
    a0 = spec.return_agent[0]
    hits = 0
    for t in range(1500):
        if t % 97 == 5:                                   # state injection and a changed clock invalidate the speculation
            st = probe._reverse_state_space[int(rs.randint(1, probe.nS))]
            spec.state = st
            plain.state = st
        if t % 131 == 7:
            spec.timestep = plain.timestep = int(rs.randint(0, 99))
        if t % 211 == 9:                                  # the caller draws from the env's generator itself
            assert spec.np_random.random() == plain.np_random.random()
        if t % 307 == 11:                                 # ... reseeds it behind the env's back
            spec.np_random.seed(t)
            plain.np_random.seed(t)
        if t == 1000:                                     # ... or replaces it
            spec.np_random = np.random.RandomState(5)
            plain.np_random = np.random.RandomState(5)
        act = {'player_a': int(rs.randint(5)), 'player_b': int(rs.randint(5))} if spec.multiagent else {a0: int(rs.randint(5))}
        before = spec._spec_seq
        out_s, out_p = spec.step(act), plain.step(act)
        assert out_s == out_p, t
        assert spec.state == plain.state and spec.timestep == plain.timestep and spec.needs_reset == plain.needs_reset
        for k in out_s[1]:                                # -0.0 == 0.0: compare the sign too (SIM:243-244, B = -A)
            assert np.signbit(out_s[1][k]) == np.signbit(out_p[1][k])
        if spec.needs_reset:
            if t % 3 == 0:
                assert spec.reset(seed=t) == plain.reset(seed=t)
            else:
                assert spec.reset() == plain.reset()
        hits += spec._spec_key is not None
    assert hits > 1300                                    # speculative launches went out for (nearly) every step
    assert spec.np_random.random() == plain.np_random.random()      # the generators end in the same state
