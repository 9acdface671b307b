"""GPU parity tests: the CUDA path (through the C ABI / its Python mirror) against
  (1) golden vectors produced by the unmodified reference (tests/golden/, oracle/make_golden.py)
  (2) the CPU oracle (oracle/soccer_oracle.c) on the same seeded inputs.
Bar: bit-exact for obs / flags / state; rewards are exactly representable (+1, -1, 0) and are
compared with ==; fp64 probabilities / Pmat / Rmat are compared with == (same operation order).
"""
import numpy as np
import pytest

from .conftest import golden_tags, load_golden, parse_tag

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _policy(g):
    return None if "policy" not in g else {s: int(a) for s, a in enumerate(g["policy"])}


def _env_kwargs(tag, g):
    w, h, slip, mode = parse_tag(tag)
    kw = dict(width=w, height=h, slip_prob=slip)
    if mode == "a_free":
        kw["player_b_policy"] = _policy(g)
    elif mode == "b_free":
        kw["player_a_policy"] = _policy(g)
    return kw, mode


# ----------------------------------------------------------------------------- K3: sweep / tables
@pytest.mark.parametrize("tag", golden_tags("table"))
def test_sweep_tables_match_reference(dev, tag):
    """env.P / env.P_readable built from the sweep kernel == the reference's (SIM:167-293):
    list lengths and order, fp64 probabilities (==), next observation, next tuple, reward, done."""
    from gym_soccer_littman94_b200.envs import SoccerSimultaneousEnv
    g = load_golden("table", tag)
    kw, mode = _env_kwargs(tag, g)
    env = SoccerSimultaneousEnv(device=dev, **kw)
    assert env.nS == int(g["nS"]) and env.width == int(g["width"])
    assert len(env.unreachable_states) == int(g["n_unreachable"])
    assert {tuple(int(v) for v in r[:5]): float(r[5]) for r in g["goal_states"]} == env.goal_states
    assert [(p, st) for p, st in env.isd] == [(float(p), tuple(int(v) for v in s))
                                              for p, s in zip(g["isd_prob"], g["isd_state"])]
    tuples = g["tuples"]
    for obs in range(1, env.nS):
        st = tuple(int(v) for v in tuples[obs])
        assert env.state_space[st] == obs and env._observation_to_state(obs) == st
    P, PR = env.P, env.P_readable
    keys = [(a, b) for a in range(5) for b in range(5)] if mode == "multi" else list(range(5))
    names = env.ACTION_STRING
    cnt, prob, nxt, rew, done, ntup = (g[k] for k in ("count", "prob", "next_obs", "reward", "done", "next_tuple"))
    assert sorted(P.keys()) == list(range(env.nS))
    for s in range(env.nS):
        assert sorted(P[s].keys()) == sorted(keys)
        for ki, k in enumerate(keys):
            tl = P[s][k]
            n = int(cnt[s, ki])
            assert len(tl) == n, (s, k)
            for j, (p, ns, r, d) in enumerate(tl):
                assert p == prob[s, ki, j] and ns == nxt[s, ki, j] and r == rew[s, ki, j] and d == bool(done[s, ki, j])
                assert isinstance(p, float) and isinstance(ns, int) and isinstance(r, float) and isinstance(d, bool)
            if s >= 1:
                kk = (names[k[0]], names[k[1]]) if mode == "multi" else names[k]
                trl = PR[tuple(int(v) for v in tuples[s])][kk]
                assert [t[1] for t in trl] == [tuple(int(v) for v in ntup[s, ki, j]) for j in range(n)]


@pytest.mark.parametrize("tag", [t for t in golden_tags("table") if "pmat_val" in load_golden("table", t)])
def test_dense_pmat_rmat_match_reference(dev, tag):
    from gym_soccer_littman94_b200.envs import SoccerSimultaneousEnv
    g = load_golden("table", tag)
    kw, _ = _env_kwargs(tag, g)
    env = SoccerSimultaneousEnv(device=dev, **kw)
    P, R = env.Pmat, env.Rmat
    assert tuple(g["pmat_shape"]) == P.shape
    nz = np.nonzero(P)
    assert np.array_equal(np.stack(nz, axis=1), g["pmat_idx"].astype(np.int64))
    assert np.array_equal(P[nz], g["pmat_val"])
    assert np.array_equal(R, g["rmat"])


# ----------------------------------------------------------------------------- K1: lock-step step
def _run_vec(dev, g, kw, kernel, n_pad=0, misalign=0):
    """Replay a golden rollout through SoccerVecEnv.  n_pad extra envs (copies of env 0) make N
    not a multiple of 4; misalign > 0 offsets every input/output pointer by that many elements."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    T, N0 = g["act_a"].shape
    N = N0 + n_pad

    def widen(a):
        return a if n_pad == 0 else np.concatenate([a, np.repeat(a[..., :1], n_pad, axis=-1)], axis=-1)
    act_a = _t(widen(g["act_a"]), dev)
    act_b = _t(widen(g["act_b"]), dev) if "act_b" in g else None
    rng8 = _t(widen(g["rng8"]), dev)
    rng32 = _t(widen(g["rng32"]).view(np.int32), dev) if "rng32" in g else None
    env = SoccerVecEnv(N, device=dev, rng_mode="injected", kernel=kernel, **kw)
    init = _t((widen(g["init_rng"]) & 3) << 2, dev)
    init_obs = env.reset(init).cpu().numpy()
    assert np.array_equal(init_obs[:N0], g["init_obs"])
    m = misalign
    obs = torch.zeros((T, N + m), dtype=torch.int32, device=dev)
    rew = torch.zeros((T, N + m), dtype=torch.float32, device=dev)
    flg = torch.zeros((T, N + m), dtype=torch.uint8, device=dev)
    rob = torch.zeros((T, N + m), dtype=torch.int32, device=dev)
    if m:
        # shift the inputs by m elements as well so that every pointer is misaligned
        def shift(x):
            buf = torch.zeros(x.numel() + m, dtype=x.dtype, device=dev)
            buf[m:] = x.reshape(-1)
            return buf
    for t in range(T):
        ins = [act_a[t], None if act_b is None else act_b[t], rng8[t], None if rng32 is None else rng32[t]]
        if m:
            ins = [None if x is None else shift(x)[m:] for x in ins]
        if env.policy_a is not None:      # b_free: the golden's single action column drives player B
            ins[0], ins[1] = None, ins[0]
        env.step(ins[0], ins[1], ins[2], rng32=ins[3], out=(obs[t, m:], rew[t, m:], flg[t, m:], rob[t, m:]))
    torch.cuda.synchronize()
    return [x[:, m:m + N0].cpu().numpy() for x in (obs, rew, flg, rob)]


@pytest.mark.parametrize("tag", golden_tags("rollout"))
def test_step_injected_matches_reference(dev, tag):
    """obs / reward / terminated / truncated / post-reset obs, step by step, against the replayed
    reference (every golden rollout: slip 0 and 0.2, folded policies, 5x4 ... 11x7)."""
    g = load_golden("rollout", tag)
    kw, _ = _env_kwargs(tag, g)
    obs, rew, flg, rob = _run_vec(dev, g, kw, "rules")
    assert np.array_equal(obs, g["obs"])
    assert np.array_equal(rew, g["reward"])
    assert np.array_equal(flg & 3, g["flags"])
    assert np.array_equal(rob, g["reset_obs"])


@pytest.mark.parametrize("tag", golden_tags("rollout"))
def test_step_dict_api_matches_reference(dev, tag):
    """The reference-shaped batched surface (reset_dict / step_dict): dict keys, reward sign of player B
    (SIM:401-402), bool dones / truncateds and info["p"] (SIM:405) against the replayed reference."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    g = load_golden("rollout", tag)
    kw, mode = _env_kwargs(tag, g)
    T, N = g["act_a"].shape
    T = min(T, 150)
    env = SoccerVecEnv(N, device=dev, rng_mode="injected", kernel="rules", **kw)
    obs0, info0 = env.reset_dict(rng8=_t((g["init_rng"] & 3) << 2, dev))
    agents = {"multi": ["player_a", "player_b"], "a_free": ["player_a"], "b_free": ["player_b"]}[mode]
    assert list(obs0) == agents and list(info0) == agents
    assert np.array_equal(obs0[agents[0]].cpu().numpy(), g["init_obs"])
    for t in range(T):
        if mode == "multi":
            action = {"player_a": _t(g["act_a"][t], dev), "player_b": _t(g["act_b"][t], dev)}
        else:
            action = {agents[0]: _t(g["act_a"][t], dev)}
        rng32 = _t(g["rng32"][t].view(np.int32), dev) if "rng32" in g else None
        obs, rew, dones, truncs, infos = env.step_dict(action, rng8=_t(g["rng8"][t], dev), rng32=rng32)
        assert list(obs) == agents and list(rew) == agents and list(dones) == agents and list(infos) == agents
        a0 = agents[0]
        assert np.array_equal(obs[a0].cpu().numpy(), g["obs"][t])
        assert np.array_equal(rew[a0].cpu().numpy(), g["reward"][t])
        if mode == "multi":
            assert np.array_equal(rew["player_b"].cpu().numpy(), -g["reward"][t])
        assert dones[a0].dtype == torch.bool and truncs[a0].dtype == torch.bool
        assert np.array_equal(dones[a0].cpu().numpy(), (g["flags"][t] & 1).astype(bool))
        assert np.array_equal(truncs[a0].cpu().numpy(), (g["flags"][t] & 2).astype(bool))
        assert np.array_equal(infos[a0]["p"].cpu().numpy(), g["info_p"][t]), t
        assert np.array_equal(env.reset_obs.cpu().numpy(), g["reset_obs"][t])
    with pytest.raises(AssertionError):
        env.step_dict([0, 1])
    with pytest.raises(AssertionError):
        env.step_dict({})


@pytest.mark.parametrize("tag", ["5x4_s000_multi", "6x4_s000_multi"])
def test_step_table_kernel_matches_reference(dev, tag):
    g = load_golden("rollout", tag)
    w, h, _, _ = parse_tag(tag)
    obs, rew, flg, rob = _run_vec(dev, g, dict(width=w, height=h, slip_prob=0.0), "table")
    assert np.array_equal(obs, g["obs"]) and np.array_equal(rew, g["reward"])
    assert np.array_equal(flg, g["flags"]) and np.array_equal(rob, g["reset_obs"])


@pytest.mark.parametrize("n_pad,misalign", [(0, 0), (1, 0), (3, 0), (0, 1), (2, 3)])
def test_step_table_slip_kernel_matches_reference(dev, n_pad, misalign):
    """slip_prob = 0.2 through the shared-memory table (soccer_step_table_slip): the 9-combination fp64
    categorical walk over the table row, against the replayed reference; vector and scalar shapes."""
    g = load_golden("rollout", "5x4_s020_multi")
    obs, rew, flg, rob = _run_vec(dev, g, dict(width=5, height=4, slip_prob=0.2), "table", n_pad, misalign)
    assert np.array_equal(obs, g["obs"]) and np.array_equal(rew, g["reward"])
    assert np.array_equal(flg, g["flags"]) and np.array_equal(rob, g["reset_obs"])


@pytest.mark.parametrize("slip", [0.2, 0.5, 1.0, 1e-9])
@pytest.mark.parametrize("draw", ["rng32", "rngf64"])
def test_step_table_slip_vs_rules_kernel_and_oracle(dev, oracle, slip, draw):
    """Slip table kernel == generic rules kernel == oracle on 20,000 envs x 60 steps for several slip values
    (1.0: the intended move has probability 0 and its combinations are skipped, SIM:226-227; 1e-9: thresholds
    next to 1.0) with 32-bit and raw fp64 draws (fp64 includes u = 0 and the largest double below 1)."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    N, T = 20000, 60
    rs = np.random.RandomState(int(slip * 1000) + len(draw))
    init = rs.randint(0, 4, N).astype(np.uint8)
    envs = {k: SoccerVecEnv(N, slip_prob=slip, device=dev, kernel=k) for k in ("rules", "table")}
    for e in envs.values():
        e.reset(_t(init << 2, dev))
    m = oracle.OracleModel(5, 4, slip)
    states = np.zeros(N, oracle.STATE_DTYPE)
    for i in range(N):
        states[i] = m.isd[int(init[i])][1]
    ts = np.zeros(N, np.int32)
    act_a, act_b = (rs.randint(0, 5, (T, N)).astype(np.uint8) for _ in range(2))
    rng8 = rs.randint(0, 16, (T, N)).astype(np.uint8)
    r32 = rs.randint(0, 2**32, (T, N), dtype=np.uint64).astype(np.uint32)
    r32[0, :4] = [0, 2**32 - 1, 2**31, 2**31 - 1]
    if draw == "rng32":
        eo, er, ef, ero = m.rollout_injected(states, ts, act_a, act_b, rng8, rng32=r32, n_threads=8)
    f64 = (r32.astype(np.float64) + 0.5) / 4294967296.0
    for t in range(T):
        outs = {}
        for k, e in envs.items():
            if draw == "rng32":
                o = e.step(_t(act_a[t], dev), _t(act_b[t], dev), _t(rng8[t], dev), rng32=_t(r32[t].view(np.int32), dev))
            else:
                u = f64[t].copy()
                if t == 1:
                    u[:3] = [0.0, np.nextafter(1.0, 0.0), 0.5]
                o = e.step(_t(act_a[t], dev), _t(act_b[t], dev), _t(rng8[t], dev), rngf64=_t(u, dev))
            outs[k] = [x.cpu().numpy().copy() for x in o]
        for x, y in zip(outs["rules"], outs["table"]):
            assert np.array_equal(x, y), (t, slip, draw)
        if draw == "rng32":
            assert np.array_equal(outs["table"][0], eo[t]) and np.array_equal(outs["table"][1], er[t])
            assert np.array_equal(outs["table"][2], ef[t]) and np.array_equal(outs["table"][3], ero[t])
    assert np.array_equal(envs["rules"].current_obs().cpu().numpy(), envs["table"].current_obs().cpu().numpy())


def test_6x4_table_kernels_vs_oracle(dev, oracle):
    """The 6x4 pitch (nS = 1105, 221 KB table: the largest that fits an SM) through every table kernel --
    K2 uniform and with table policies, K1 slip, the fused replay -- against the oracle."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    N, K, seed = 4100, 90, 77
    m = oracle.OracleModel(6, 4, 0.0)
    rs = np.random.RandomState(11)
    for pol in (False, True):
        pa = rs.randint(0, 5, m.nS).astype(np.int8) if pol else None
        env = SoccerVecEnv(N, width=6, height=4, device=dev, rng_mode="philox", kernel="table", seed=seed)
        states, ts = m.states_from_obs(env.reset().cpu().numpy()), np.zeros(N, np.int32)
        eo, er, ef, es = m.rollout_philox(states, ts, K, seed, policy_a=pa, n_threads=8)
        obs, rew, flg, stats = env.rollout(K, policy_a=pa)
        assert np.array_equal(obs.cpu().numpy(), eo) and np.array_equal(rew.cpu().numpy(), er)
        assert np.array_equal(flg.cpu().numpy(), ef) and np.array_equal(stats.cpu().numpy(), es)
    # injected: fused replay (slip 0) and slip table step (slip 0.3)
    T = 40
    init = rs.randint(0, 4, N).astype(np.uint8)
    act_a, act_b = (rs.randint(0, 5, (T, N)).astype(np.uint8) for _ in range(2))
    rng8 = rs.randint(0, 16, (T, N)).astype(np.uint8)
    r32 = rs.randint(0, 2**32, (T, N), dtype=np.uint64).astype(np.uint32)
    for slip in (0.0, 0.3):
        mm = oracle.OracleModel(6, 4, slip)
        states = np.zeros(N, oracle.STATE_DTYPE)
        for i in range(N):
            states[i] = mm.isd[int(init[i])][1]
        ts = np.zeros(N, np.int32)
        eo, er, ef, ero = mm.rollout_injected(states, ts, act_a, act_b, rng8, rng32=r32 if slip else None, n_threads=8)
        env = SoccerVecEnv(N, width=6, height=4, slip_prob=slip, device=dev, kernel="table")
        env.reset(_t(init << 2, dev))
        if slip == 0.0:
            obs, rew, flg, rob = env.step_many(_t(act_a, dev), _t(act_b, dev), _t(rng8, dev))
            got = [x.cpu().numpy() for x in (obs, rew, flg, rob)]
        else:
            rows = [[x.cpu().numpy().copy() for x in env.step(_t(act_a[t], dev), _t(act_b[t], dev), _t(rng8[t], dev),
                                                               rng32=_t(r32[t].view(np.int32), dev))] for t in range(T)]
            got = [np.stack([r[j] for r in rows]) for j in range(4)]
        for x, y in zip(got, (eo, er, ef, ero)):
            assert np.array_equal(x, y), slip


def test_table_env_generic_options_roundtrip(dev):
    """A kernel='table' env still offers the generic kernel's options (Philox draws, detail flags): the step
    runs on a CELL-layout copy of the INDEX-layout state and converts back."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    N = 4100
    et = SoccerVecEnv(N, device=dev, kernel="table", rng_mode="philox", seed=3)
    er = SoccerVecEnv(N, device=dev, kernel="rules", rng_mode="philox", seed=3)
    assert torch.equal(et.reset(), er.reset())
    g = torch.Generator(device=dev).manual_seed(0)
    for t in range(40):
        a, b = (torch.randint(0, 5, (N,), dtype=torch.uint8, device=dev, generator=g) for _ in range(2))
        ot, orr = et.step(a, b, detail=True), er.step(a, b, detail=True)
        for x, y in zip(ot, orr):
            assert torch.equal(x, y), t
    assert torch.equal(et.current_obs(), er.current_obs()) and torch.equal(et.timesteps(), er.timesteps())


@pytest.mark.parametrize("kernel", ["rules", "table"])
@pytest.mark.parametrize("n_pad,misalign", [(1, 0), (3, 0), (0, 1), (2, 3)])
def test_step_ragged_and_misaligned(dev, kernel, n_pad, misalign):
    """N not a multiple of 4 and pointers off the 16-byte grid take the scalar path: same answers."""
    g = load_golden("rollout", "5x4_s000_multi")
    sub = {k: (g[k][:300] if g[k].ndim == 2 else g[k]) for k in g.files}
    obs, rew, flg, rob = _run_vec(dev, sub, dict(width=5, height=4, slip_prob=0.0), kernel, n_pad, misalign)
    assert np.array_equal(obs, sub["obs"]) and np.array_equal(rew, sub["reward"])
    assert np.array_equal(flg & 3, sub["flags"]) and np.array_equal(rob, sub["reset_obs"])


@pytest.mark.parametrize("w,h,slip", [(14, 9, 0.0), (18, 7, 0.0), (31, 4, 0.0), (9, 6, 0.3)])
def test_maximum_pitch_sizes_vs_oracle(dev, oracle, w, h, slip):
    """The largest supported pitches (width*height = 126 field cells, nS = 31,501; odd and even
    heights): K1 (byte-parallel and generic), K2 and the sweep against the oracle."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    m = oracle.OracleModel(w, h, slip)
    assert m.nS == 1 + 2 * (w * h) * (w * h - 1)
    N, T = 257, 160
    rs = np.random.RandomState(w * 100 + h)
    # start from random reachable states so that the whole pitch is exercised
    obs0 = rs.randint(1, m.nS, N).astype(np.int32)
    t0 = rs.randint(0, 99, N).astype(np.int32)
    states, ts = m.states_from_obs(obs0), t0.copy()
    act_a, act_b = (rs.randint(0, 5, (T, N)).astype(np.uint8) for _ in range(2))
    rng8 = rs.randint(0, 16, (T, N)).astype(np.uint8)
    rng32 = rs.randint(0, 2 ** 32, (T, N), dtype=np.uint64).astype(np.uint32) if slip else None
    eo, er, ef, ero = m.rollout_injected(states, ts, act_a, act_b, rng8, rng32, n_threads=4)
    env = SoccerVecEnv(N, width=w, height=h, slip_prob=slip, device=dev, kernel="rules")
    env.set_state(_t(obs0, dev), _t(t0, dev))
    for t in range(T):
        o, r, f, ro = env.step(_t(act_a[t], dev), _t(act_b[t], dev), _t(rng8[t], dev),
                               rng32=None if rng32 is None else _t(rng32[t].view(np.int32), dev))
        assert np.array_equal(o.cpu().numpy(), eo[t]) and np.array_equal(r.cpu().numpy(), er[t]), t
        assert np.array_equal(f.cpu().numpy() & 3, ef[t]) and np.array_equal(ro.cpu().numpy(), ero[t]), t
    assert np.array_equal(env.current_obs().cpu().numpy(), m.obs_from_states(states))
    # K2 with Philox draws
    e2 = SoccerVecEnv(N - 1, width=w, height=h, slip_prob=slip, device=dev, rng_mode="philox", seed=9)
    st2 = m.states_from_obs(e2.reset().cpu().numpy())
    po, pr, pf, ps = m.rollout_philox(st2, np.zeros(N - 1, np.int32), 130, 9, n_threads=4)
    o2, r2, f2, s2 = e2.rollout(130)
    assert np.array_equal(o2.cpu().numpy(), po) and np.array_equal(f2.cpu().numpy(), pf)
    assert np.array_equal(s2.cpu().numpy(), ps)


def test_step_empty_batch(dev):
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    env = SoccerVecEnv(0, device=dev, kernel="rules")
    e = torch.zeros(0, dtype=torch.uint8, device=dev)
    env.reset(e)
    env.step(e, e, e)
    env.rollout(4)
    torch.cuda.synchronize()


@pytest.mark.parametrize("kernel", ["table", "rules"])
@pytest.mark.parametrize("n,chunks", [(1000, 1), (4099, 3), (70001, 8)])
def test_step_host_equals_device_path(dev, kernel, n, chunks):
    """The host-buffer C entry point (pipelined chunks, wide and narrow downloads) returns exactly
    what the device-resident step returns, step after step (state carried across calls)."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    rs = np.random.RandomState(n)
    ref = SoccerVecEnv(n, device=dev, kernel=kernel)
    envs = {(nr, zc): SoccerVecEnv(n, device=dev, kernel=kernel) for nr in (False, True) for zc in (False, True)}
    init = rs.randint(0, 16, n).astype(np.uint8)
    for e in [ref] + list(envs.values()):
        e.reset(_t(init, dev))
    for t in range(60):
        a, b, r = (rs.randint(0, hi, n).astype(np.uint8) for hi in (5, 5, 16))
        obs, rew, flg, _ = ref.step(_t(a, dev), _t(b, dev), _t(r, dev))
        ha, hb, hr = (torch.from_numpy(x).pin_memory() for x in (a, b, r))
        for (nr, zc), e in envs.items():
            # zero_copy: the kernel (wide or fused-narrow outputs) reads / writes the pinned host buffers itself
            ho, hw, hf = e.step_host(ha, hb, hr, narrow=nr, n_chunks=chunks, zero_copy=zc)
            ho = ho.numpy().view(np.uint16) if nr else ho.numpy()
            assert np.array_equal(ho.astype(np.int32), obs.cpu().numpy()), (t, nr, zc)
            assert np.array_equal(hw.numpy().astype(np.float32), rew.cpu().numpy())
            assert np.array_equal(hf.numpy(), flg.cpu().numpy())
    for e in envs.values():
        assert torch.equal(e.state, ref.state)


@pytest.mark.parametrize("w,h", [(5, 4), (6, 4)])
@pytest.mark.parametrize("n,chunks", [(1000, 1), (4099, 3), (70001, 8), (1 << 18, 4)])
def test_packed_step_equals_plain_step(dev, w, h, n, chunks):
    """soccer_step_table_packed (joint-action byte in, one 16-bit result word out) on device tensors, through the
    staged host path and through the zero-copy host path == soccer_step_table, step after step, states included;
    the ragged tail (n % 4 != 0) goes through the scalar kernel."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    rs = np.random.RandomState(n + w)
    ref = SoccerVecEnv(n, width=w, height=h, device=dev, kernel="table")
    envs = [SoccerVecEnv(n, width=w, height=h, device=dev, kernel="table") for _ in range(3)]
    init = rs.randint(0, 16, n).astype(np.uint8)
    for e in [ref] + envs:
        e.reset(_t(init, dev))
    for t in range(110):                                   # beyond the 100-step truncation
        a, b, r = (rs.randint(0, hi, n).astype(np.uint8) for hi in (5, 5, 16))
        if t % 7 == 3:
            a[:] = 0; b[:] = 0                              # NOOP stretches let episodes reach the truncation
        obs, rew, flg, _ = ref.step(_t(a, dev), _t(b, dev), _t(r, dev))
        j = SoccerVecEnv.pack_joint(torch.from_numpy(a), torch.from_numpy(b))
        assert np.array_equal(j.numpy(), a | (b << 4))
        hj, hr = j.pin_memory(), torch.from_numpy(r).pin_memory()
        words = [envs[0].step_packed(hj.to(dev), hr.to(dev)).cpu(),
                 envs[1].step_host_packed(hj, hr, n_chunks=chunks, zero_copy=False).clone(),
                 envs[2].step_host_packed(hj, hr, zero_copy=True).clone()]
        for wd in words:
            o, rw, term, trunc = SoccerVecEnv.unpack_result(wd)
            assert np.array_equal(o.numpy(), obs.cpu().numpy()), t
            assert np.array_equal(rw.numpy(), rew.cpu().numpy())
            assert np.array_equal((term.numpy() * 1 + trunc.numpy() * 2).astype(np.uint8), flg.cpu().numpy())
    for e in envs:
        assert torch.equal(e.state, ref.state)


def test_host_arena(dev):
    """soccer_host_alloc: pinned, device-mapped, 4 KB-aligned tensors that copy both ways and are released with
    the last tensor carved from them."""
    import gc
    from gym_soccer_littman94_b200 import _lib
    ar = _lib.HostArena(3 << 20)
    a, b = ar.take(1000, torch.uint8), ar.take(70000, torch.int16)
    assert a.is_pinned() and b.is_pinned() and a.data_ptr() % 4096 == 0 and b.data_ptr() % 4096 == 0
    a.copy_(torch.arange(1000) % 251); b.copy_(torch.arange(70000) % 30000)
    da, db = a.to(dev, non_blocking=True), b.to(dev, non_blocking=True)
    torch.cuda.synchronize()
    assert int(da.sum()) == int((torch.arange(1000) % 251).sum()) and int(db.to(torch.int64).sum()) == int((torch.arange(70000) % 30000).sum())
    b.copy_(db + 1, non_blocking=True)
    torch.cuda.synchronize()
    assert int(b[69999]) == 69999 % 30000 + 1
    with pytest.raises(_lib.SoccerB200Error):
        ar.take(4 << 20, torch.uint8)
    del ar, a, b
    gc.collect()
    env_bufs = __import__("gym_soccer_littman94_b200.envs", fromlist=["SoccerVecEnv"]).SoccerVecEnv(4099, device=dev).alloc_host_inputs()
    assert len(env_bufs) == 3 and all(t.is_pinned() and t.numel() == 4099 and t.dtype == torch.uint8 for t in env_bufs)


def test_packed_step_rejects_what_it_cannot_do(dev):
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    z = torch.zeros(8, dtype=torch.uint8)
    with pytest.raises(NotImplementedError):
        SoccerVecEnv(8, device=dev, kernel="rules").step_host_packed(z, z)
    with pytest.raises(NotImplementedError):
        SoccerVecEnv(8, slip_prob=0.2, device=dev, kernel="table").step_packed(z.to(dev), z.to(dev))
    with pytest.raises(ValueError):
        SoccerVecEnv(8, device=dev, kernel="table").step_host_packed(z[:4], z)


@pytest.mark.parametrize("n,off", [(0, 0), (5, 0), (4096, 0), (100003, 0), (70000, 3)])
def test_step_stats_kernel(dev, n, off):
    """soccer_step_stats == the same counts taken with numpy (vector path, tail, misaligned)."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    env = SoccerVecEnv(8, device=dev, kernel="rules")
    rs = np.random.RandomState(n + off)
    f = rs.choice([0, 0, 0, 1, 2, 3], size=n + off).astype(np.uint8)
    r = np.where(f & 1, rs.choice([-1.0, 1.0], size=n + off), 0.0).astype(np.float32)
    st = torch.zeros(6, dtype=torch.int64, device=dev)
    st[5] = 77
    for _ in range(2):
        env.step_stats(flags=_t(f, dev)[off:], reward=_t(r, dev)[off:], stats=st)
    f, r = f[off:], r[off:]
    want = 2 * np.array([(f != 0).sum(), (r > 0).sum(), (r < 0).sum(), (f == 2).sum(), n, 0]) + [0, 0, 0, 0, 0, 77]
    assert np.array_equal(st.cpu().numpy(), want)


@pytest.mark.parametrize("kernel", ["rules", "table"])
def test_config2_4096_envs_vs_oracle(dev, oracle, kernel):
    """BASELINE config 2: 4096 lock-step envs, host-supplied joint actions, injected draws,
    T = 10,000 steps, bit-exact against the oracle (itself pinned to the reference)."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    N, T, CH = 4096, 10000, 500
    rs = np.random.RandomState(2024)
    m = oracle.OracleModel(5, 4, 0.0)
    init = rs.randint(0, 4, N).astype(np.uint8)
    states = np.zeros(N, oracle.STATE_DTYPE)
    for i in range(N):
        states[i] = m.isd[int(init[i])][1]
    ts = np.zeros(N, np.int32)
    env = SoccerVecEnv(N, device=dev, kernel=kernel)
    env.reset(_t(init << 2, dev))
    n_eps = 0
    for c in range(T // CH):
        act_a = rs.randint(0, 5, (CH, N)).astype(np.uint8)
        act_b = rs.randint(0, 5, (CH, N)).astype(np.uint8)
        rng8 = rs.randint(0, 16, (CH, N)).astype(np.uint8)
        eo, er, ef, ero = m.rollout_injected(states, ts, act_a, act_b, rng8, n_threads=8)
        da, db, dr = _t(act_a, dev), _t(act_b, dev), _t(rng8, dev)
        obs = torch.empty((CH, N), dtype=torch.int32, device=dev)
        rew = torch.empty((CH, N), dtype=torch.float32, device=dev)
        flg = torch.empty((CH, N), dtype=torch.uint8, device=dev)
        rob = torch.empty((CH, N), dtype=torch.int32, device=dev)
        if c % 2 == 0:                      # alternate the per-step entry point and the T-step one
            for t in range(CH):
                env.step(da[t], db[t], dr[t], out=(obs[t], rew[t], flg[t], rob[t]))
        else:
            env.step_many(da, db, dr, out=(obs, rew, flg, rob))
        assert np.array_equal(obs.cpu().numpy(), eo), f"obs mismatch in chunk {c}"
        assert np.array_equal(rew.cpu().numpy(), er)
        assert np.array_equal(flg.cpu().numpy(), ef)
        assert np.array_equal(rob.cpu().numpy(), ero)
        n_eps += int((ef != 0).sum())
    assert np.array_equal(env.current_obs().cpu().numpy(), m.obs_from_states(states))
    assert np.array_equal(env.timesteps().cpu().numpy(), ts)
    assert n_eps > 1_000_000   # ~34-step episodes: the fused reset really is exercised


@pytest.mark.parametrize("kernel", ["rules", "table"])
@pytest.mark.parametrize("n,T,want_ro", [(1, 1, True), (5, 3, True), (4096, 2, False), (4096, 37, True), (4099, 9, True),
                                         (100000, 4, False), (33 * 148 * 4 + 8, 70, True), (0, 5, True)])
def test_step_many_fused_equals_single_steps(dev, kernel, n, T, want_ro):
    """soccer_step_many runs the T steps in ONE launch with the state in registers (k_replay*): same
    streams and final state as T launches of K1, for vector / scalar shapes, T not a multiple of the
    register buffer depth, T shorter than it, partial last pass, with and without reset_obs."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    g = torch.Generator(device=dev).manual_seed(n * 131 + T)
    a, b, r = (torch.randint(0, hi, (T, n), dtype=torch.uint8, device=dev, generator=g) for hi in (5, 5, 16))
    init = torch.randint(0, 16, (n,), dtype=torch.uint8, device=dev, generator=g)
    e1 = SoccerVecEnv(n, device=dev, kernel=kernel, want_reset_obs=want_ro)
    e2 = SoccerVecEnv(n, device=dev, kernel=kernel, want_reset_obs=want_ro)
    e1.reset(init)
    e2.reset(init)
    # start mid-episode so that truncation (t >= 100) fires inside the window as well
    warm = 80
    for t in range(warm):
        for e in (e1, e2):
            e.step(a[t % T], b[(t * 7) % T], r[(t * 3) % T])
    assert torch.equal(e1.state, e2.state)
    want = [tuple(None if x is None else x.clone() for x in e1.step(a[t], b[t], r[t])) for t in range(T)]
    obs, rew, flg, rob = e2.step_many(a, b, r)
    for t in range(T):
        assert torch.equal(obs[t], want[t][0]), (t, "obs")
        assert torch.equal(rew[t], want[t][1]), (t, "reward")
        assert torch.equal(flg[t], want[t][2]), (t, "flags")
        if want_ro:
            assert torch.equal(rob[t], want[t][3]), (t, "reset_obs")
        else:
            assert rob is None
    assert torch.equal(e1.state, e2.state)
    if n >= 4096 and T >= 30:
        assert int((flg & 2).count_nonzero()) > 0 and int((flg & 1).count_nonzero()) > 0


def test_exhaustive_single_steps_vs_golden(dev):
    """Every state x 25 joint actions x 4 draw values in ONE launch of each kernel, against the
    reference's table: the sweep of BASELINE config 5 done through the step entry points."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    g = load_golden("table", "5x4_s000_multi")
    nS = int(g["nS"])
    s, ja, r = np.meshgrid(np.arange(1, nS), np.arange(25), np.arange(4), indexing="ij")
    s, ja, r = s.ravel(), ja.ravel(), r.ravel()
    cnt = g["count"][s, ja].astype(np.int64)
    slot = np.where(cnt == 4, r, np.where(cnt == 2, r >> 1, 0))
    want_obs = g["next_obs"][s, ja, slot]
    want_rew = g["reward"][s, ja, slot].astype(np.float32)
    want_done = g["done"][s, ja, slot]
    N = len(s)
    for kernel in ("rules", "table"):
        env = SoccerVecEnv(N, device=dev, kernel=kernel)
        env.set_state(_t(s.astype(np.int32), dev))
        obs, rew, flg, rob = env.step(_t((ja // 5).astype(np.uint8), dev), _t((ja % 5).astype(np.uint8), dev),
                                      _t(r.astype(np.uint8), dev))
        assert np.array_equal(obs.cpu().numpy(), want_obs), kernel
        assert np.array_equal(rew.cpu().numpy(), want_rew), kernel
        assert np.array_equal(flg.cpu().numpy() & 1, want_done), kernel
        # fused reset: terminated envs restart from isd[0] (reset draw bits are 0 here)
        ro = rob.cpu().numpy()
        assert np.array_equal(ro[want_done == 1], np.full(int(want_done.sum()), int(g["isd_obs"][0])))
        assert np.array_equal(ro[want_done == 0], want_obs[want_done == 0])


# ----------------------------------------------------------------------------- K2 / Philox
@pytest.mark.parametrize("kernel", ["rules", "table"])
def test_rollout_philox_vs_oracle(dev, oracle, kernel):
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    N, K, seed, base = 2048, 150, 1234567, 10_000_000_000
    m = oracle.OracleModel(5, 4, 0.0)
    env = SoccerVecEnv(N, device=dev, rng_mode="philox", kernel=kernel, seed=seed, env_id_base=base)
    init_obs = env.reset().cpu().numpy()
    # initial reset draw = reset bits of word(seed, env, 2^64-1)
    want0 = np.array([m.state_to_obs(m.isd[oracle.philox_decode(oracle.philox_word(seed, base + i, (1 << 64) - 1))[3]][1])
                      for i in range(N)], np.int32)
    assert np.array_equal(init_obs, want0)
    states = m.states_from_obs(init_obs)
    ts = np.zeros(N, np.int32)
    # two launches with an unaligned step0 in between (K = 150 then 61)
    for kk in (K, 61):
        step0 = env.step_count
        eo, er, ef, es = m.rollout_philox(states, ts, kk, seed, step0=step0, env_id_base=base, n_threads=8)
        obs, rew, flg, stats = env.rollout(kk)
        assert np.array_equal(obs.cpu().numpy(), eo)
        assert np.array_equal(rew.cpu().numpy(), er)
        assert np.array_equal(flg.cpu().numpy(), ef)
        assert np.array_equal(stats.cpu().numpy(), es)
    assert np.array_equal(env.current_obs().cpu().numpy(), m.obs_from_states(states))


@pytest.mark.parametrize("kernel", ["rules", "table"])
@pytest.mark.parametrize("N,which", [(1023, "ab"), (1024, "a"), (2048, "b"), (77, "ab")])
def test_rollout_table_policy_and_ragged(dev, oracle, kernel, N, which):
    """Table policies for one or both players (utils/policies.py format), vector and N % 4 != 0
    shapes, through the rules K2 and the shared-memory-table K2 (policies ride next to the table)."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    K, seed = 120, 99
    m = oracle.OracleModel(5, 4, 0.0)
    rs = np.random.RandomState(5)
    pa = rs.randint(0, 5, m.nS).astype(np.int8) if "a" in which else None
    pb = rs.randint(0, 5, m.nS).astype(np.int8) if "b" in which else None
    env = SoccerVecEnv(N, device=dev, rng_mode="philox", kernel=kernel, seed=seed)
    states = m.states_from_obs(env.reset().cpu().numpy())
    ts = np.zeros(N, np.int32)
    eo, er, ef, es = m.rollout_philox(states, ts, K, seed, policy_a=pa, policy_b=pb, n_threads=4)
    obs, rew, flg, stats = env.rollout(K, policy_a=pa, policy_b=pb)
    assert np.array_equal(obs.cpu().numpy(), eo) and np.array_equal(rew.cpu().numpy(), er)
    assert np.array_equal(flg.cpu().numpy(), ef) and np.array_equal(stats.cpu().numpy(), es)
    assert np.array_equal(env.current_obs().cpu().numpy(), m.obs_from_states(states))


@pytest.mark.parametrize("w,h,n,kernel", [(5, 4, 516, "rules"), (7, 5, 203, "rules"), (5, 4, 516, "table"), (5, 4, 4099, "table")])
def test_rollout_and_step_philox_with_slip_vs_oracle(dev, oracle, w, h, n, kernel):
    """slip_prob = 0.2 (the commented registration default, gym_soccer/__init__.py:8) through K2 and
    through K1 in Philox mode: the word's 32-bit step draw, fp64 cumulative sums in the reference's order."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    K, seed = 70, 4242
    m = oracle.OracleModel(w, h, 0.2)
    env = SoccerVecEnv(n, width=w, height=h, slip_prob=0.2, device=dev, rng_mode="philox", seed=seed, kernel=kernel)
    init = env.reset().cpu().numpy()
    states, ts = m.states_from_obs(init), np.zeros(n, np.int32)
    eo, er, ef, es = m.rollout_philox(states, ts, K, seed, n_threads=8)
    obs, rew, flg, stats = env.rollout(K)
    assert np.array_equal(obs.cpu().numpy(), eo) and np.array_equal(rew.cpu().numpy(), er)
    assert np.array_equal(flg.cpu().numpy(), ef) and np.array_equal(stats.cpu().numpy(), es)
    assert (ef != 0).sum() > 50
    # the same trajectory step by step through K1 (Philox draws, caller-supplied = Philox-decoded actions)
    if n > 1000:
        return
    e1 = SoccerVecEnv(n, width=w, height=h, slip_prob=0.2, device=dev, rng_mode="philox", seed=seed, kernel=kernel)
    e1.reset()
    for k in range(K):
        acts = np.array([oracle.philox_decode(oracle.philox_word(seed, i, k))[:2] for i in range(n)], np.uint8)
        o, r, f, _ = e1.step(_t(acts[:, 0].copy(), dev), _t(acts[:, 1].copy(), dev))
        assert np.array_equal(o.cpu().numpy(), eo[k]) and np.array_equal(f.cpu().numpy() & 3, ef[k]), k


@pytest.mark.parametrize("kernel,N,base", [("rules", 512, 0), ("table", 512, 0), ("table", 4099, 1 << 33), ("rules", 4099, 77)])
def test_step_philox_equals_rollout(dev, kernel, N, base):
    """K1 in Philox mode (soccer_step_philox / soccer_step_table_philox, ragged tail included) fed the actions K2
    draws for itself follows the same trajectory."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    from oracle import soccer_oracle as so
    K, seed = 40, 31337
    a = SoccerVecEnv(N, device=dev, rng_mode="philox", kernel="rules", seed=seed, env_id_base=base)
    b = SoccerVecEnv(N, device=dev, rng_mode="philox", kernel=kernel, seed=seed, env_id_base=base)
    a.reset(); b.reset()
    obs, rew, flg, _ = a.rollout(K)
    for k in range(K):
        acts = np.array([so.philox_decode(so.philox_word(seed, base + i, k))[:2] for i in range(N)], np.uint8)
        o, r, f, ro = b.step(_t(acts[:, 0].copy(), dev), _t(acts[:, 1].copy(), dev))
        assert torch.equal(o, obs[k]) and torch.equal(r, rew[k]) and torch.equal(f, flg[k])
        assert torch.equal(ro, b.current_obs())


def test_rollout_gpu_count_independence(dev):
    """Sharding contract: envs [0, N) stepped as one batch == two half batches with env_id_base."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    N, K, seed = 4096, 64, 7
    full = SoccerVecEnv(N, device=dev, rng_mode="philox", seed=seed)
    full.reset()
    fo, fr, ff, fs = full.rollout(K)
    tot = torch.zeros(6, dtype=torch.int64, device=dev)
    for half in range(2):
        e = SoccerVecEnv(N // 2, device=dev, rng_mode="philox", seed=seed, env_id_base=half * N // 2)
        e.reset()
        o, r, f, s = e.rollout(K)
        sl = slice(half * N // 2, (half + 1) * N // 2)
        assert torch.equal(o, fo[:, sl]) and torch.equal(r, fr[:, sl]) and torch.equal(f, ff[:, sl])
        tot += s
    assert torch.equal(tot, fs)


# (the distributional check of Philox-mode play against the reference lives in tests/test_distribution.py: chi-square /
#  z-tests against the exact distribution derived from the reference's Pmat, no literal statistics)


# ----------------------------------------------------------------------------- full-size properties
def test_full_size_properties_2_24(dev):
    """BASELINE's bandwidth configuration (2^24 envs): the two K1 kernels agree bit for bit, and
    size-independent invariants hold (reward != 0 iff terminated; obs == 0 iff terminated; reset_obs
    is a start state wherever flags != 0; timestep resets)."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    N = 1 << 24
    g = torch.Generator(device=dev).manual_seed(0)
    outs = {}
    for kernel in ("rules", "table"):
        env = SoccerVecEnv(N, device=dev, kernel=kernel)
        g.manual_seed(0)
        env.reset(torch.randint(0, 16, (N,), dtype=torch.uint8, device=dev, generator=g))
        acc = torch.zeros(4, dtype=torch.int64, device=dev)
        for t in range(40):
            a = torch.randint(0, 5, (N,), dtype=torch.uint8, device=dev, generator=g)
            b = torch.randint(0, 5, (N,), dtype=torch.uint8, device=dev, generator=g)
            r = torch.randint(0, 16, (N,), dtype=torch.uint8, device=dev, generator=g)
            obs, rew, flg, rob = env.step(a, b, r)
            term = (flg & 1) != 0
            assert torch.equal(term, rew != 0) and torch.equal(term, obs == 0)
            ended = flg != 0
            starts = torch.tensor([253, 254, 435, 436], device=dev, dtype=torch.int32)
            assert torch.isin(rob[ended], starts).all()
            assert torch.equal(rob[~ended], obs[~ended])
            assert torch.equal(env.current_obs(), rob)
            assert (env.timesteps()[ended] == 0).all()
            acc += torch.stack([obs.sum(dtype=torch.int64), (rew > 0).sum(), (rew < 0).sum(), rob.sum(dtype=torch.int64)])
        outs[kernel] = acc.cpu().numpy()
        del env
    assert np.array_equal(outs["rules"], outs["table"])


class _Guarded:
    """Output tensors carved out of larger buffers pre-filled with a sentinel byte: after the kernels ran, every
    byte outside the carved views must still be the sentinel (compute-sanitizer is not available on the pool, so
    this is the out-of-bounds-write check).  `skew` shifts a view by that many ELEMENTS to force misaligned paths."""
    PAD = 256  # bytes on either side

    def __init__(self, dev):
        self.dev, self.bufs = dev, []

    def make(self, shape, dtype, skew=0):
        numel = int(np.prod(shape))
        es = torch.empty(0, dtype=dtype).element_size()
        raw = torch.full((2 * self.PAD + (numel + skew) * es,), 0xA5, dtype=torch.uint8, device=self.dev)
        lo = self.PAD + skew * es
        view = raw[lo:lo + numel * es].view(dtype).view(shape)
        self.bufs.append((raw, lo, numel * es))
        return view

    def check(self):
        torch.cuda.synchronize()
        for raw, lo, nb in self.bufs:
            assert bool((raw[:lo] == 0xA5).all()) and bool((raw[lo + nb:] == 0xA5).all()), "write outside an output buffer"


@pytest.mark.parametrize("kernel", ["rules", "table"])
@pytest.mark.parametrize("n,skew", [(7, 0), (1000, 0), (4099, 0), (4099, 1), (65541, 0), (65536, 3)])
def test_no_writes_outside_the_output_buffers(dev, kernel, n, skew):
    """Every K1 / K2 / replay / statistics entry point, ragged and misaligned batches, outputs and the state tensor
    inside sentinel-filled guard bands."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    i32, f32, u8, i16, i8 = torch.int32, torch.float32, torch.uint8, torch.int16, torch.int8
    g = _Guarded(dev)
    rs = np.random.RandomState(n + skew)
    a, b, r = (_t(rs.randint(0, hi, n).astype(np.uint8), dev) for hi in (5, 5, 16))
    T = 5
    A, B, R = (_t(rs.randint(0, hi, (T, n)).astype(np.uint8), dev) for hi in (5, 5, 16))

    def fresh(**kw):
        e = SoccerVecEnv(n, device=dev, kernel=kernel, **kw)
        st = g.make((n,), i32, skew)
        st.copy_(e.state)
        e.state = st
        return e
    # K1 injected (+ reset_obs), K1 Philox, K1 narrow / packed on device tensors
    e = fresh()
    e.reset(r)
    e.step(a, b, r, out=(g.make((n,), i32, skew), g.make((n,), f32, skew), g.make((n,), u8, skew), g.make((n,), i32, skew)))
    e.step(a, b, r, out=(g.make((n,), i32, skew), g.make((n,), f32, skew), g.make((n,), u8, skew), None))
    e.step_many(A, B, R, out=(g.make((T, n), i32, skew), g.make((T, n), f32, skew), g.make((T, n), u8, skew), None))
    st6 = g.make((6,), torch.int64)
    st6.zero_()
    e.step_stats(stats=st6)
    if kernel == "table":
        e.step_packed(a | (b << 4), r, out=g.make((n,), i16, skew))
    from gym_soccer_littman94_b200 import _lib
    import ctypes as C
    o16, r8, f8 = g.make((n,), i16, skew), g.make((n,), i8, skew), g.make((n,), u8, skew)
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(e.lib.soccer_step_narrow(C.byref(e.pitch), None if e.table is None else p(e.table), p(e.state), p(a), p(b),
                                        p(r), p(o16), p(r8), p(f8), n, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
               "soccer_step_narrow")
    ep = fresh(rng_mode="philox", seed=3)
    ep.reset()
    ep.step(a, b, out=(g.make((n,), i32, skew), g.make((n,), f32, skew), g.make((n,), u8, skew), g.make((n,), i32, skew)))
    # K2 with streams and statistics
    ep.rollout(T, out=(g.make((T, n), i32, skew), g.make((T, n), f32, skew), g.make((T, n), u8, skew)), stats=st6)
    g.check()


@pytest.mark.parametrize("w,h,kernel", [(5, 4, "table"), (5, 4, "rules"), (7, 5, "rules")])
@pytest.mark.parametrize("n,skew", [(1001, 0), (4100, 1)])
def test_no_writes_outside_the_output_buffers_slip(dev, w, h, kernel, n, skew):
    """The slip_prob > 0 steppers (table walk, rules walk, K2 SLIP instantiations) inside guard bands."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    i32, f32, u8 = torch.int32, torch.float32, torch.uint8
    g = _Guarded(dev)
    rs = np.random.RandomState(n)
    a, b, r = (_t(rs.randint(0, hi, n).astype(np.uint8), dev) for hi in (5, 5, 16))
    r32 = _t(rs.randint(-2**31, 2**31 - 1, n).astype(np.int32), dev)
    for mode in ("injected", "philox"):
        e = SoccerVecEnv(n, width=w, height=h, slip_prob=0.2, device=dev, kernel=kernel, rng_mode=mode, seed=5)
        st = g.make((n,), i32, skew)
        st.copy_(e.state)
        e.state = st
        e.reset(r if mode == "injected" else None)
        out = (g.make((n,), i32, skew), g.make((n,), f32, skew), g.make((n,), u8, skew), g.make((n,), i32, skew))
        if mode == "injected":
            e.step(a, b, r, rng32=r32, out=out)
        else:
            e.step(a, b, out=out)
            e.rollout(3, out=(g.make((3, n), i32, skew), g.make((3, n), f32, skew), g.make((3, n), u8, skew)))
    g.check()


@pytest.mark.parametrize("k_save,k_load", [("table", "table"), ("rules", "table"), ("table", "rules")])
def test_checkpoint_resume_continues_the_trajectory(dev, k_save, k_load, tmp_path):
    """state_dict / load_state_dict: 2 x 48 steps with a save / torch.save / load into a FRESH env in between (also
    across kernels: the checkpoint is layout-neutral) == 96 steps in one go; timesteps and statistics included."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    N, seed = 4099, 77
    ref = SoccerVecEnv(N, device=dev, rng_mode="philox", kernel="rules", seed=seed, env_id_base=1000)
    ref.reset()
    ro, rr, rf, rs = ref.rollout(96)
    a = SoccerVecEnv(N, device=dev, rng_mode="philox", kernel=k_save, seed=seed, env_id_base=1000)
    a.reset()
    o1, r1, f1, s1 = a.rollout(48)
    torch.save(a.state_dict(), tmp_path / "ckpt.pt")
    b = SoccerVecEnv(N, device=dev, rng_mode="philox", kernel=k_load, seed=0, env_id_base=0)
    b.load_state_dict(torch.load(tmp_path / "ckpt.pt"))
    o2, r2, f2, s2 = b.rollout(48)
    assert torch.equal(torch.cat([o1, o2]), ro) and torch.equal(torch.cat([r1, r2]), rr) and torch.equal(torch.cat([f1, f2]), rf)
    assert torch.equal(s1 + s2, rs)
    assert torch.equal(b.current_obs(), ref.current_obs()) and torch.equal(b.timesteps(), ref.timesteps())
    with pytest.raises(ValueError):
        SoccerVecEnv(N + 1, device=dev).load_state_dict(a.state_dict())


@pytest.mark.parametrize("slip", [0.2, 0.5, 1.0, 1e-9])
@pytest.mark.parametrize("n", [4099, 1 << 16, (1 << 18) + 4])
def test_slip_fast_path_with_queue_equals_walk(dev, slip, n, monkeypatch):
    """soccer_step_table_slip with the slip index (constant-prefix fast path, deferred envs walked from a shared-memory
    queue) == the in-place walk of every env, step after step: obs, reward, flags, reset_obs and states; uint32 and
    raw fp64 draws incl. u = 0, u just below 1 and u on the constant thresholds; a population forced onto collision
    states so that whole warps are deferred."""
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    rs = np.random.RandomState(n)
    envs = {}
    for mode in ("walk", "queued"):
        monkeypatch.setenv("SOCCER_B200_SLIP_WALK", "1" if mode == "walk" else "0")
        envs[mode] = SoccerVecEnv(n, slip_prob=slip, device=dev, kernel="table")
    assert envs["queued"].slip_index is not None
    init = _t(rs.randint(0, 16, n).astype(np.uint8), dev)
    for e in envs.values():
        e.reset(init)
    # E_k of slip_prob (sequential fp64 sums of the 9 combination probabilities, SIM:209-223 order)
    sp = slip
    mp = [(1 - sp) ** 2] + [(1 - sp) * sp / 2] * 4 + [sp * sp / 4] * 4
    E = np.cumsum(np.array(mp, np.float64))
    for t in range(60):
        a, b, r = (_t(rs.randint(0, hi, n).astype(np.uint8), dev) for hi in (5, 5, 16))
        if t == 20:                                            # adjacent players walking into each other: collisions everywhere
            obs_adj = envs["walk"].current_obs().clone()
            from gym_soccer_littman94_b200.envs import SoccerSimultaneousEnv
            idx = SoccerSimultaneousEnv(device=dev)._state_to_observation((1, 2, 1, 3, 0))
            obs_adj[::2] = idx
            for e in envs.values():
                e.set_state(obs_adj)
            a[:] = 3; b[:] = 4
        outs = {}
        if t % 2 == 0:
            r32 = _t(rs.randint(-2**31, 2**31 - 1, n).astype(np.int32), dev)
            for mode, e in envs.items():
                monkeypatch.setenv("SOCCER_B200_SLIP_WALK", "1" if mode == "walk" else "0")
                outs[mode] = [x.clone() for x in e.step(a, b, r, rng32=r32)]
        else:
            u = rs.random_sample(n)
            u[:9] = E; u[9:18] = np.nextafter(E, 0); u[18] = 0.0; u[19] = 1.0 - 2.0 ** -53; u[20:29] = np.nextafter(E, 2)
            uf = _t(u, dev)
            for mode, e in envs.items():
                monkeypatch.setenv("SOCCER_B200_SLIP_WALK", "1" if mode == "walk" else "0")
                outs[mode] = [x.clone() for x in e.step(a, b, r, rngf64=uf)]
        for x, y in zip(outs["walk"], outs["queued"]):
            assert torch.equal(x, y), t
        assert torch.equal(envs["walk"].state, envs["queued"].state), t
