"""Planners over the transition dynamics, on the GPU.

Same entry points, arguments and return tuples as the reference's gym_soccer/utils/planners.py (= PL) for the
single-agent (folded-policy) env they are written for.

* value_iteration (PL:4-18), policy_evaluation (PL:20-31), policy_improvement (PL:33-43) and policy_iteration
  (PL:45-55) walk `env.P` lists in the reference.  Here they call the hand-written kernels behind
  `soccer_plan` / `soccer_bellman_q` (csrc/soccer_planner.cuh): one thread per (observation, action) enumerates
  the same list in the same order and accumulates with the reference's fp64 operation order, and a whole value
  iteration or policy evaluation runs inside ONE cooperative launch.  Results are bit-identical to the
  reference's: V, Q, the greedy policy and the number of sweeps (tests/test_gpu_planners.py compares with ==).
* policy_eval (PL:57-70) and modified_policy_iteration (PL:73-87) are written against the dense `Pmat` / `Rmat`
  in the reference (np.dot).  Pmat has at most 15 non-zeros per (state, action), so here the contraction runs
  sparsely over the same on-the-fly transition lists, in the hand-written kernels `soccer_dense_q` and
  `soccer_policy_eval` (a whole policy_eval loop in one cooperative launch) -- no dense matrix, no library GEMM.
  Rmat comes out bit-identical; Pmat . v is summed in list order where the reference's BLAS has its own order, so
  these agree to fp64 round-off (|V - V_ref| < 1e-9, identical greedy policies).
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib

MAX_SWEEPS = 1 << 30        # the reference loops until convergence; this only guards against theta <= 0


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _single_agent(env):
    assert not env.multiagent, "planners act on a single-agent env (one player's policy folded, SIM:266-279)"
    return env._lib, C.byref(env._pitch), _ptr(env._pol_a), _ptr(env._pol_b), env.device


def _plan(env, pi_in, theta, gamma):
    """soccer_plan: value iteration (pi_in None) or policy evaluation, one cooperative kernel launch."""
    lib, pitch, pa, pb, dev = _single_agent(env)
    nS, nA = env.nS, env.nA
    nbytes = C.c_int64()
    _lib.check(lib.soccer_plan_workspace_bytes_host(pitch, C.byref(nbytes)), "soccer_plan_workspace_bytes_host")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    V = torch.empty(nS, dtype=torch.float64, device=dev)
    Q = torch.empty((nS, nA), dtype=torch.float64, device=dev) if pi_in is None else None
    pi = torch.empty(nS, dtype=torch.int32, device=dev) if pi_in is None else None
    sweeps = torch.zeros(1, dtype=torch.int32, device=dev)
    pin = None if pi_in is None else torch.as_tensor(np.asarray(pi_in), dtype=torch.int32, device=dev).contiguous()
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.soccer_plan(pitch, pa, pb, _ptr(pin), float(theta), float(gamma), MAX_SWEEPS, _ptr(V), _ptr(Q),
                                   _ptr(pi), _ptr(sweeps), _ptr(ws), st), "soccer_plan")
    return V, Q, pi, int(sweeps.item())


def _backup_q(env, V, gamma):
    """soccer_bellman_q: Q[nS, nA] of one backup of V (a device or host vector)."""
    lib, pitch, pa, pb, dev = _single_agent(env)
    Vd = torch.as_tensor(np.asarray(V) if not torch.is_tensor(V) else V, dtype=torch.float64, device=dev).contiguous()
    Q = torch.empty((env.nS, env.nA), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.soccer_bellman_q(pitch, pa, pb, _ptr(Vd), float(gamma), _ptr(Q), st), "soccer_bellman_q")
    return Q


def _dense_q(env, v, gamma):
    """soccer_dense_q: q[nS, nA] = Rmat + gamma * (Pmat . v), contracted sparsely over the transition lists (PL:78-79)."""
    lib, pitch, pa, pb, dev = _single_agent(env)
    q = torch.empty((env.nS, env.nA), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.soccer_dense_q(pitch, pa, pb, _ptr(v), float(gamma), _ptr(q), st), "soccer_dense_q")
    return q


def _policy_eval(env, policy, theta, gamma, k, init):
    """soccer_policy_eval: the whole PL:57-70 loop in one cooperative launch; device tensors in and out."""
    lib, pitch, pa, pb, dev = _single_agent(env)
    nbytes = C.c_int64()
    _lib.check(lib.soccer_policy_eval_workspace_bytes_host(pitch, env.nA, C.byref(nbytes)), "soccer_policy_eval_workspace_bytes_host")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    V = torch.empty(env.nS, dtype=torch.float64, device=dev)
    sweeps = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.soccer_policy_eval(pitch, pa, pb, _ptr(policy), _ptr(init), float(theta), float(gamma),
                                          int(min(k, MAX_SWEEPS)), _ptr(V), _ptr(sweeps), _ptr(ws), st), "soccer_policy_eval")
    return V, sweeps


def value_iteration(env, theta, discount_factor):
    """PL:4-18.  Returns (pi, V, Q, sweeps); V is the value BEFORE the last sweep, as there."""
    V, Q, pi, cc = _plan(env, None, theta, discount_factor)
    return pi.cpu().numpy().astype(np.int64), V.cpu().numpy(), Q.cpu().numpy(), cc


def policy_evaluation(pi, env, theta, discount_factor):
    """PL:20-31: iterate V[s] <- sum over P[s][pi[s]] from zero until the sup-norm change < theta."""
    V, _, _, _ = _plan(env, pi, theta, discount_factor)
    return V.cpu().numpy()


def policy_improvement(V, env, discount_factor):
    """PL:33-43."""
    Q = _backup_q(env, V, discount_factor).cpu().numpy()
    return np.argmax(Q, axis=1), Q


def policy_iteration(env, theta, discount_factor):
    """PL:45-55 (random initial policy from numpy's global generator, as there)."""
    cc = 0
    pi = np.random.choice(tuple(range(env.nA)), env.nS)
    while True:
        old_pi = pi.copy()
        V = policy_evaluation(pi, env, theta, discount_factor)
        pi, Q = policy_improvement(V, env, discount_factor)
        cc += 1
        if np.all(old_pi == pi):
            break
    return pi, V, Q, cc


def policy_eval(env, policy, theta, discount_factor, k=10000000, init=None):
    """planners.py:57-70: at most k sweeps of v <- r_pi + gamma P_pi v for a stochastic policy [nS, nA]; returns (v, sweeps).
    Like the reference, a given `init` array receives the result (PL:58, 67: `v[:] = value_fc`)."""
    dev = env.device
    pol = torch.as_tensor(np.asarray(policy), dtype=torch.float64, device=dev).contiguous()
    assert tuple(pol.shape) == (env.nS, env.nA), "policy must be [nS, nA]"
    v0 = None if init is None else torch.as_tensor(np.asarray(init), dtype=torch.float64, device=dev).contiguous()
    V, sweeps = _policy_eval(env, pol, theta, discount_factor, k, v0)
    out = V.cpu().numpy()
    if isinstance(init, np.ndarray):
        init[:] = out
        out = init
    return out, int(sweeps.item())


def modified_policy_iteration(env, k, theta, discount_factor):
    """planners.py:73-87.  Returns (pi, greedy_v, q, outer_iterations).  The dense contractions of the reference
    (np.dot over Pmat / Rmat) run on the hand-written sparse kernels soccer_dense_q / soccer_policy_eval."""
    dev, nA = env.device, env.nA
    v = torch.zeros(env.nS, dtype=torch.float64, device=dev)
    threshold = (theta * (1 - discount_factor)) / (2 * discount_factor)
    counter = 0
    while True:
        q = _dense_q(env, v, discount_factor)
        greedy_v, best = q.max(dim=1)                      # np.max / np.argmax (first maximum)
        best = torch.argmax((q == greedy_v[:, None]).to(torch.uint8), dim=1)
        if float((v - greedy_v).abs().max()) <= threshold:
            return best.cpu().numpy(), greedy_v.cpu().numpy(), q.cpu().numpy(), counter
        policy = torch.zeros((env.nS, nA), dtype=torch.float64, device=dev)
        policy[torch.arange(env.nS, device=dev), best] = 1.0          # np.eye(nA)[best_action]
        v, _ = _policy_eval(env, policy, theta, discount_factor, k, greedy_v)
        counter += 1
