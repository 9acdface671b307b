"""Planners over the transition table, on the GPU.

Same entry points, arguments and return tuples as the reference's gym_soccer/utils/planners.py
(value_iteration :4-18, policy_iteration :45-55, modified_policy_iteration :73-87), for the
single-agent (folded-policy) env they are written for.  Where the reference walks Python lists of
(prob, next_state, reward, done) per (s, a), these run batched Bellman backups
        Q = Rmat + gamma * Pmat . V          (one fp64 contraction over next states per sweep)
on the dense Pmat[nS, nS, nA] / Rmat[nS, nA] that the `soccer_dense` kernel emits in the reference's
own accumulation order.  The `* (not done)` factor of the reference (planners.py:11) only ever
multiplies V[0] -- done transitions lead to the terminal observation 0 -- and V[0] stays 0, so it is
dropped.  Results agree with the reference to fp64 round-off (the summation order over next states
differs): identical greedy policies, |V - V_ref| < 1e-9 (tests/test_gpu_planners.py).
"""
import numpy as np
import torch


def _dense(env):
    """(Pmat[nS, nS, nA], Rmat[nS, nA]) as fp64 CUDA tensors, cached on the env."""
    cache = getattr(env, "_planner_dense", None)
    if cache is None:
        assert not env.multiagent, "planners act on a single-agent env (one player's policy folded, SIM:266-279)"
        dev = getattr(env, "device", torch.device("cuda"))
        cache = (torch.from_numpy(np.ascontiguousarray(env.Pmat)).to(dev), torch.from_numpy(np.ascontiguousarray(env.Rmat)).to(dev))
        env._planner_dense = cache
    return cache


def _backup(P, R, V, gamma):
    return R + gamma * torch.einsum("sna,n->sa", P, V)


def value_iteration(env, theta, discount_factor):
    """planners.py:4-18.  Returns (pi, V, Q, sweeps); V is the value BEFORE the last sweep, as there."""
    P, R = _dense(env)
    V = torch.zeros(P.shape[0], dtype=torch.float64, device=P.device)
    cc = 0
    while True:
        Q = _backup(P, R, V, discount_factor)
        cc += 1
        newV = Q.max(dim=1).values
        if float((V - newV).abs().max()) < theta:
            break
        V = newV
    return Q.argmax(dim=1).cpu().numpy(), V.cpu().numpy(), Q.cpu().numpy(), cc


def policy_evaluation(pi, env, theta, discount_factor):
    """planners.py:20-31: iterate V <- R_pi + gamma P_pi V from zero until the sup-norm change < theta."""
    P, R = _dense(env)
    idx = torch.as_tensor(np.asarray(pi), dtype=torch.int64, device=P.device)
    ar = torch.arange(P.shape[0], device=P.device)
    P_pi = P[ar, :, idx]                      # [nS, nS]
    R_pi = R[ar, idx]
    prev = torch.zeros_like(R_pi)
    while True:
        V = R_pi + discount_factor * (P_pi @ prev)
        if float((prev - V).abs().max()) < theta:
            break
        prev = V
    return V.cpu().numpy()


def policy_improvement(V, env, discount_factor):
    """planners.py:33-43."""
    P, R = _dense(env)
    Q = _backup(P, R, torch.as_tensor(V, dtype=torch.float64, device=P.device), discount_factor)
    return Q.argmax(dim=1).cpu().numpy(), Q.cpu().numpy()


def policy_iteration(env, theta, discount_factor):
    """planners.py:45-55 (random initial policy from numpy's global generator, as there)."""
    P, _ = _dense(env)
    cc = 0
    pi = np.random.choice(P.shape[2], P.shape[0])
    while True:
        old_pi = pi.copy()
        V = policy_evaluation(pi, env, theta, discount_factor)
        pi, Q = policy_improvement(V, env, discount_factor)
        cc += 1
        if np.all(old_pi == pi):
            break
    return pi, V, Q, cc


def policy_eval(env, policy, theta, discount_factor, k=10000000, init=None):
    """planners.py:57-70: at most k sweeps of v <- r_pi + gamma P_pi v for a stochastic policy [nS, nA]."""
    P, R = _dense(env)
    pol = torch.as_tensor(np.asarray(policy), dtype=torch.float64, device=P.device)
    v = torch.zeros(P.shape[0], dtype=torch.float64, device=P.device) if init is None \
        else torch.as_tensor(np.asarray(init), dtype=torch.float64, device=P.device).clone()
    r_pi = (pol * R).sum(dim=1)
    P_pi = torch.einsum("sna,sa->sn", P, pol)
    cc = 0
    for _ in range(k):
        new_v = r_pi + discount_factor * (P_pi @ v)
        delta = float((new_v - v).abs().max())
        v = new_v
        cc += 1
        if delta < theta:
            break
    return v.cpu().numpy(), cc


def modified_policy_iteration(env, k, theta, discount_factor):
    """planners.py:73-87.  Returns (pi, greedy_v, q, outer_iterations)."""
    P, R = _dense(env)
    nA = P.shape[2]
    v = torch.zeros(P.shape[0], dtype=torch.float64, device=P.device)
    threshold = (theta * (1 - discount_factor)) / (2 * discount_factor)
    counter = 0
    while True:
        q = _backup(P, R, v, discount_factor)
        greedy_v, best = q.max(dim=1)
        if float((v - greedy_v).abs().max()) <= threshold:
            return best.cpu().numpy(), greedy_v.cpu().numpy(), q.cpu().numpy(), counter
        policy = torch.nn.functional.one_hot(best, nA).to(torch.float64)
        v_np, _ = policy_eval(env, policy.cpu().numpy(), theta=theta, discount_factor=discount_factor, k=k,
                              init=greedy_v.cpu().numpy())
        v = torch.as_tensor(v_np, dtype=torch.float64, device=P.device)
        counter += 1
