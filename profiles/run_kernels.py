"""Short driver for ncu captures: a few launches of each hot kernel at its bench size.

    python profiles/run_kernels.py [k1_table|k1_table_stats|k1_single_agent|k1_rules|k2_table|k2_rules|k1_slip|k2_slip|k1_philox|k1_packed|rules_modes|replay|all] [--envs N]
"""
import argparse
import contextlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_soccer_littman94_b200.envs import SoccerVecEnv  # noqa: E402


@contextlib.contextmanager
def profiled():
    """cudaProfilerStart / Stop around the launches that matter: run ncu with `--profile-from-start off` and only these
    are captured (the play-in steps before them are not)."""
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    try:
        yield
    finally:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()


def k1(kernel, n, iters, with_stats=False, policy=False):
    """K1 on a played-in population (64 steps first); with_stats: the bench's headline variant (statistics fused);
    policy: single-agent mode, player B folded."""
    import numpy as np
    dev = torch.device("cuda", 0)
    kw = dict(player_b_policy=np.random.RandomState(0).randint(0, 5, 761).astype(np.int8)) if policy else {}
    env = SoccerVecEnv(n, device=dev, kernel=kernel, want_reset_obs=False, **kw)
    g = torch.Generator(device=dev).manual_seed(0)
    ins = [tuple(torch.randint(0, hi, (n,), dtype=torch.uint8, device=dev, generator=g) for hi in (5, 5, 16)) for _ in range(4)]
    env.reset(ins[0][2])
    stats = torch.zeros(6, dtype=torch.int64, device=dev) if with_stats else None
    for i in range(64):
        a, b, r = ins[i % 4]
        env.step(a, None if policy else b, r, stats=stats)
    with profiled():
        for i in range(iters):
            a, b, r = ins[i % 4]
            env.step(a, None if policy else b, r, stats=stats)


def k1_philox(n, iters):
    dev = torch.device("cuda", 0)
    env = SoccerVecEnv(n, device=dev, kernel="table", rng_mode="philox", want_reset_obs=False)
    g = torch.Generator(device=dev).manual_seed(0)
    a, b = (torch.randint(0, 5, (n,), dtype=torch.uint8, device=dev, generator=g) for _ in range(2))
    env.reset()
    for _ in range(40):
        env.step(a, b)
    with profiled():
        for _ in range(iters):
            env.step(a, b)


def k1_packed(n, iters):
    """the packed K1 step on device tensors and on pinned host buffers (zero copy: PCIe-bound)"""
    dev = torch.device("cuda", 0)
    env = SoccerVecEnv(n, device=dev, kernel="table", want_reset_obs=False)
    hj, hr = env.alloc_host_inputs(packed=True)
    hj.copy_(torch.randint(0, 5, (n,), dtype=torch.uint8) | (torch.randint(0, 5, (n,), dtype=torch.uint8) << 4))
    hr.copy_(torch.randint(0, 16, (n,), dtype=torch.uint8))
    dj, dr = hj.to(dev), hr.to(dev)
    env.reset(dr)
    ha, hb, hr3 = env.alloc_host_inputs()
    ha.copy_(hj & 15); hb.copy_(hj >> 4); hr3.copy_(hr)
    with profiled():
        for _ in range(iters):
            env.step_packed(dj, dr)
        for _ in range(iters):
            env.step_host_packed(hj, hr)
        for _ in range(iters):
            env.step_host(ha, hb, hr3, narrow=True, zero_copy=True)


def k2(kernel, n, K, iters):
    dev = torch.device("cuda", 0)
    env = SoccerVecEnv(n, device=dev, kernel=kernel, rng_mode="philox", seed=0)
    env.reset()
    bufs = (torch.empty((K, n), dtype=torch.int32, device=dev), torch.empty((K, n), dtype=torch.float32, device=dev),
            torch.empty((K, n), dtype=torch.uint8, device=dev))
    env.rollout(K, out=bufs)
    with profiled():
        for _ in range(iters):
            env.rollout(K, out=bufs)


def k1_slip(n, iters):
    """K1 with slip_prob = 0.2 on a played-in population: injected rng32 draws and Philox draws (integer fast path),
    injected fp64 draws (constant-prefix fast path + queue)."""
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    a, b, r = (torch.randint(0, hi, (n,), dtype=torch.uint8, device=dev, generator=g) for hi in (5, 5, 16))
    r32 = torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device=dev, generator=g)
    r64 = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    for mode in ("rng32", "philox", "rngf64"):
        env = SoccerVecEnv(n, slip_prob=0.2, device=dev, kernel="table", want_reset_obs=False,
                           rng_mode="philox" if mode == "philox" else "injected")
        env.reset(None if mode == "philox" else r)
        def one():
            if mode == "philox":
                env.step(a, b)
            elif mode == "rng32":
                env.step(a, b, r, rng32=r32)
            else:
                env.step(a, b, r, rngf64=r64)
        for _ in range(40):
            one()
        with profiled():
            for _ in range(iters):
                one()
        del env


def k2_slip(n, K, iters):
    dev = torch.device("cuda", 0)
    e2 = SoccerVecEnv(n, slip_prob=0.2, device=dev, kernel="table", rng_mode="philox")
    e2.reset()
    e2.rollout(64, want_streams=False)          # a played-in population
    bufs = (torch.empty((K, n), dtype=torch.int32, device=dev), torch.empty((K, n), dtype=torch.float32, device=dev),
            torch.empty((K, n), dtype=torch.uint8, device=dev))
    e2.rollout(K, out=bufs)
    with profiled():
        for _ in range(iters):
            e2.rollout(K, out=bufs)


def rules_modes(n, iters):
    """The rules kernels of a pitch without a table (7x5) in the modes round 2 moved onto the byte-parallel paths:
    K1 slip 0.2 (rng32 / Philox), K1 single-agent (slip 0 and 0.2), K2 slip 0.2, K2 with a table policy."""
    import numpy as np
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    a, b, r = (torch.randint(0, hi, (n,), dtype=torch.uint8, device=dev, generator=g) for hi in (5, 5, 16))
    r32 = torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device=dev, generator=g)
    probe = SoccerVecEnv(4, width=7, height=5, device=dev, kernel="rules")
    pol = np.random.RandomState(0).randint(0, 5, probe.nS).astype(np.int8)
    cases = [(dict(slip_prob=0.2), lambda e: e.step(a, b, r, rng32=r32)),
             (dict(slip_prob=0.2, rng_mode="philox"), lambda e: e.step(a, b)),
             (dict(player_b_policy=pol), lambda e: e.step(a, None, r)),
             (dict(player_b_policy=pol, slip_prob=0.2), lambda e: e.step(a, None, r, rng32=r32))]
    for kw, one in cases:
        env = SoccerVecEnv(n, width=7, height=5, device=dev, kernel="rules", want_reset_obs=False, **kw)
        env.reset(None if env.rng_mode == "philox" else r)
        for _ in range(30):
            one(env)
        with profiled():
            for _ in range(iters):
                one(env)
        del env
    n2, K = 1 << 22, 16
    bufs = (torch.empty((K, n2), dtype=torch.int32, device=dev), torch.empty((K, n2), dtype=torch.float32, device=dev),
            torch.empty((K, n2), dtype=torch.uint8, device=dev))
    tp = torch.from_numpy(pol).to(dev)
    for kw, rk in ((dict(slip_prob=0.2), {}), (dict(), dict(policy_a=tp)), (dict(slip_prob=0.2), dict(policy_a=tp))):
        e2 = SoccerVecEnv(n2, width=7, height=5, device=dev, kernel="rules", rng_mode="philox", **kw)
        e2.reset()
        e2.rollout(64, want_streams=False, **rk)
        with profiled():
            for _ in range(iters):
                e2.rollout(K, out=bufs, **rk)
        del e2


def replay(kernel, n, T, iters):
    dev = torch.device("cuda", 0)
    env = SoccerVecEnv(n, device=dev, kernel=kernel, want_reset_obs=False)
    a, b, r = (torch.randint(0, hi, (T, n), dtype=torch.uint8, device=dev) for hi in (5, 5, 16))
    out = (torch.empty((T, n), dtype=torch.int32, device=dev), torch.empty((T, n), dtype=torch.float32, device=dev),
           torch.empty((T, n), dtype=torch.uint8, device=dev), None)
    env.reset(r[0])
    env.step_many(a, b, r, out=out)
    with profiled():
        for _ in range(iters):
            env.step_many(a, b, r, out=out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="all")
    ap.add_argument("--envs", type=int, default=1 << 24)
    ap.add_argument("--iters", type=int, default=3)
    args = ap.parse_args()
    if args.what in ("k1_table", "all"):
        k1("table", args.envs, args.iters)
    if args.what in ("k1_table_stats", "all"):
        k1("table", args.envs, args.iters, with_stats=True)
    if args.what in ("k1_single_agent", "all"):
        k1("table", args.envs, args.iters, policy=True)
    if args.what in ("k1_rules", "all"):
        k1("rules", args.envs, args.iters)
    if args.what in ("k2_table", "all"):
        k2("table", 1 << 20, 64, args.iters)
    if args.what in ("k2_rules", "all"):
        k2("rules", 1 << 20, 64, args.iters)
    if args.what in ("k1_slip", "all"):
        k1_slip(args.envs, args.iters)
    if args.what in ("k2_slip", "all"):
        k2_slip(1 << 22, 16, args.iters)
    if args.what in ("k1_philox", "all"):
        k1_philox(args.envs, args.iters)
    if args.what in ("k1_packed", "all"):
        k1_packed(args.envs, args.iters)
    if args.what in ("rules_modes", "all"):
        rules_modes(args.envs, 2)
    if args.what in ("replay", "all"):
        replay("table", 1 << 22, 64, 3)
        replay("table", 4096, 4000, 3)
    print("done")
