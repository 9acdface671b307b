"""GPU planners (gym_soccer_littman94_b200/utils/planners.py) against the reference's planners
(tests/golden/ref_planner_*.npz, produced by oracle/make_golden.py --planner-only from
/root/reference/gym_soccer/utils/planners.py = PL), and the best-response checks the reference's own
integration tests make (/root/reference/gym_soccer/tests/test_general.py:304-458).

The list-walking planners (value_iteration, policy_evaluation, policy_improvement, policy_iteration;
PL:4-55) run on the hand-written Bellman kernels (soccer_plan / soccer_bellman_q), which reproduce the
reference's list order and fp64 operation order: V, Q, greedy policy and sweep count are compared with ==.
modified_policy_iteration (PL:73-87) is np.dot over Pmat / Rmat in the reference (BLAS order): compared to
fp64 round-off, |V - V_ref| <= 1e-9, identical greedy policies wherever the reference's top-two Q gap > 1e-8."""
import glob
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TAGS = sorted(os.path.basename(p)[len("ref_planner_"):-4] for p in glob.glob(os.path.join(GOLDEN, "ref_planner_*.npz")))


def _env(tag, g):
    from gym_soccer_littman94_b200.envs import SoccerSimultaneousEnv
    slip = int(tag.split("_")[1][1:]) / 100.0
    pol = {s: int(a) for s, a in enumerate(g["policy"])}
    kw = {"player_b_policy": pol} if tag.endswith("a_free") else {"player_a_policy": pol}
    return SoccerSimultaneousEnv(width=5, height=4, slip_prob=slip, **kw)


def _same_policy(pi, ref_pi, ref_Q):
    top2 = np.sort(ref_Q, axis=1)[:, -2:]
    decided = (top2[:, 1] - top2[:, 0]) > 1e-8
    return np.array_equal(pi[decided], ref_pi[decided])


@pytest.mark.skipif(not TAGS, reason="planner goldens not generated")
@pytest.mark.parametrize("tag", TAGS)
def test_planners_match_reference(tag):
    from gym_soccer_littman94_b200.utils import planners
    g = np.load(os.path.join(GOLDEN, f"ref_planner_{tag}.npz"))
    env = _env(tag, g)
    theta, gamma = float(g["theta"]), float(g["gamma"])
    pi, V, Q, cc = planners.value_iteration(env, theta, gamma)
    assert cc == int(g["vi_cc"])                                            # same number of sweeps
    assert np.array_equal(V, g["vi_V"]) and np.array_equal(Q, g["vi_Q"])    # bit for bit
    assert np.array_equal(pi, g["vi_pi"])
    # policy evaluation of a fixed policy, one improvement step, full policy iteration (PL:20-55)
    pe_V = planners.policy_evaluation(g["pe_pi"], env, float(g["pe_theta"]), gamma)
    assert np.array_equal(pe_V, g["pe_V"])
    imp_pi, imp_Q = planners.policy_improvement(g["pe_V"], env, gamma)
    assert np.array_equal(imp_Q, g["imp_Q"]) and np.array_equal(imp_pi, g["imp_pi"])
    np.random.seed(int(g["it_seed"]))
    ppi, pV, pQ, pcc = planners.policy_iteration(env, float(g["pe_theta"]), gamma)
    assert pcc == int(g["it_cc"]) and np.array_equal(ppi, g["it_pi"])
    assert np.array_equal(pV, g["it_V"]) and np.array_equal(pQ, g["it_Q"])
    # dense planners: fp64 round-off
    mpi, mV, mQ, mcc = planners.modified_policy_iteration(env, 1, theta, gamma)
    assert np.abs(mV - g["mpi_V"]).max() <= 1e-9 and _same_policy(mpi, g["mpi_pi"], g["mpi_Q"])


def test_bellman_q_joint_actions_vs_dense():
    """soccer_bellman_q on the multi-agent env gives the joint-action backup Q[s, aa, ab] (what minimax-Q needs):
    checked against Rmat + gamma * Pmat . V from the dense kernel's (reference-pinned) matrices."""
    import ctypes as C
    from gym_soccer_littman94_b200 import _lib
    from gym_soccer_littman94_b200.envs import SoccerSimultaneousEnv
    env = SoccerSimultaneousEnv(width=5, height=4, slip_prob=0.2)
    dev = env.device
    V = torch.rand(env.nS, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    V[0] = 0.0
    Q = torch.empty((env.nS, 25), dtype=torch.float64, device=dev)
    _lib.check(env._lib.soccer_bellman_q(C.byref(env._pitch), None, None, C.c_void_p(V.data_ptr()), 0.9,
                                         C.c_void_p(Q.data_ptr()), None), "soccer_bellman_q")
    torch.cuda.synchronize()
    P = torch.from_numpy(np.ascontiguousarray(env.Pmat)).to(dev)       # [nS, nS, 5, 5]
    R = torch.from_numpy(np.ascontiguousarray(env.Rmat)).to(dev)       # [nS, 5, 5]
    want = (R + 0.9 * torch.einsum("snab,n->sab", P, V)).reshape(env.nS, 25)
    want[0] = 0.0           # P[0]: the absorbing terminal observation, every entry done (Pmat[0, 0] holds the quirk sum)
    assert float((Q - want).abs().max()) < 1e-12


def test_best_response_beats_stand_and_random_policies():
    """test_general.py:304-458: the value-iteration best response wins 100 % against the stand
    policy and > 95 % against a random policy, for either side -- played here on the GPU through the
    fused rollout kernel with both table policies on the device."""
    from gym_soccer_littman94_b200.envs import SoccerSimultaneousEnv, SoccerVecEnv
    from gym_soccer_littman94_b200.utils import planners
    from gym_soccer_littman94_b200.utils.policies import get_random_policy, get_stand_policy, policy_to_table
    # the reference asserts == 1.0 over 1,000 episodes; over ~50,000 episodes with slip_prob 0.2 a
    # handful of truncations may appear, hence 0.998
    for opponent, min_win in ((get_stand_policy(761), 0.998), (get_random_policy(761, 5, seed=42), 0.95)):
        for side in ("a", "b"):
            kw = {"player_b_policy": opponent} if side == "a" else {"player_a_policy": opponent}
            env = SoccerSimultaneousEnv(width=5, height=4, slip_prob=0.2, **kw)
            pi, V, Q, cc = planners.value_iteration(env, 1e-10, 0.99)
            br, opp = pi.astype(np.int8), policy_to_table(opponent, 761)
            vec = SoccerVecEnv(4096, slip_prob=0.2, device=env.device, rng_mode="philox", kernel="rules", seed=3)
            vec.reset()
            pa, pb = (br, opp) if side == "a" else (opp, br)
            _, _, _, st = vec.rollout(400, policy_a=pa, policy_b=pb, want_streams=False)
            ep, ga, gb, tr, steps, length = [int(x) for x in st.cpu().numpy()]
            wins = ga if side == "a" else gb
            assert ep > 10000 and wins / ep >= min_win, (side, wins, ep, tr)


@pytest.mark.parametrize("side,slip", [("a", 0.2), ("b", 0.0)])
def test_dense_planners_vs_numpy_on_the_dense_matrices(side, slip):
    """policy_eval / the q backup of modified_policy_iteration (PL:57-87) on the hand-written SPARSE kernels
    (soccer_policy_eval, soccer_dense_q) against the reference's own formulas evaluated with numpy on the dense
    Pmat / Rmat (which test_dense_pmat_rmat_match_reference pins bit for bit to the reference): every sweep count,
    values to fp64 round-off (the reference sums with BLAS, the kernel in list order)."""
    from gym_soccer_littman94_b200.envs import SoccerSimultaneousEnv
    from gym_soccer_littman94_b200.utils import planners
    from gym_soccer_littman94_b200.utils.policies import get_random_policy
    opp = get_random_policy(761, 5, seed=7)
    kw = {"player_b_policy": opp} if side == "a" else {"player_a_policy": opp}
    env = SoccerSimultaneousEnv(width=5, height=4, slip_prob=slip, **kw)
    P, R = np.asarray(env.Pmat), np.asarray(env.Rmat)                   # [nS, nS, 5], [nS, 5]
    rs = np.random.RandomState(3)
    policy = rs.dirichlet(np.ones(5), size=env.nS)
    gamma = 0.97

    def ref_policy_eval(theta, k, init):                                 # PL:57-70, vectorised over s
        v = np.zeros(env.nS) if init is None else init.copy()
        cc = 0
        for _ in range(k):
            r_pi = (policy * R).sum(axis=1)
            pv = np.einsum("sna,n->sa", P, v)
            val = r_pi + gamma * (pv * policy).sum(axis=1)
            delta = np.abs(val - v).max()
            v = val
            cc += 1
            if delta < theta:
                break
        return v, cc
    for theta, k, init in ((1e-10, 10000000, None), (1e-3, 10000000, rs.rand(env.nS)), (0.0, 7, rs.rand(env.nS)), (1.0, 0, rs.rand(env.nS))):
        if init is not None:
            init[0] = 0.0          # (Pmat[0, 0] holds the reference's quirk sum of 160: v[0] != 0 would blow up there as here)
        want, wcc = ref_policy_eval(theta, k, init)
        got, cc = planners.policy_eval(env, policy, theta, gamma, k=k, init=None if init is None else init.copy())
        assert cc == wcc, (theta, k)
        assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
    v = rs.rand(env.nS)
    q = planners._dense_q(env, torch.as_tensor(v, device=env.device), gamma).cpu().numpy()
    assert np.abs(q - (R + gamma * np.einsum("sna,n->sa", P, v))).max() <= 1e-12 * 160
    # the quirk row: Pmat[0, 0, a] = number of goal states x sum of the combination probabilities
    assert np.allclose(q[0], gamma * P[0, 0] * v[0], rtol=1e-15)
