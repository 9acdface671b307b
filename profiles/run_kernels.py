"""Short driver for ncu captures: a few launches of each hot kernel at its bench size.

    python profiles/run_kernels.py [k1_table|k1_rules|k2_table|k2_rules|k1_slip|k2_slip|k1_philox|k1_packed|replay|all] [--envs N]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_soccer_littman94_b200.envs import SoccerVecEnv  # noqa: E402


def k1(kernel, n, iters):
    dev = torch.device("cuda", 0)
    env = SoccerVecEnv(n, device=dev, kernel=kernel, want_reset_obs=False)
    g = torch.Generator(device=dev).manual_seed(0)
    a, b, r = (torch.randint(0, hi, (n,), dtype=torch.uint8, device=dev, generator=g) for hi in (5, 5, 16))
    env.reset(r)
    for _ in range(iters):
        env.step(a, b, r)
    torch.cuda.synchronize()


def k1_philox(n, iters):
    dev = torch.device("cuda", 0)
    env = SoccerVecEnv(n, device=dev, kernel="table", rng_mode="philox", want_reset_obs=False)
    g = torch.Generator(device=dev).manual_seed(0)
    a, b = (torch.randint(0, 5, (n,), dtype=torch.uint8, device=dev, generator=g) for _ in range(2))
    env.reset()
    for _ in range(iters):
        env.step(a, b)
    torch.cuda.synchronize()


def k1_packed(n, iters):
    """the packed K1 step on device tensors and on pinned host buffers (zero copy: PCIe-bound)"""
    dev = torch.device("cuda", 0)
    env = SoccerVecEnv(n, device=dev, kernel="table", want_reset_obs=False)
    hj, hr = env.alloc_host_inputs(packed=True)
    hj.copy_(torch.randint(0, 5, (n,), dtype=torch.uint8) | (torch.randint(0, 5, (n,), dtype=torch.uint8) << 4))
    hr.copy_(torch.randint(0, 16, (n,), dtype=torch.uint8))
    dj, dr = hj.to(dev), hr.to(dev)
    env.reset(dr)
    for _ in range(iters):
        env.step_packed(dj, dr)
    for _ in range(iters):
        env.step_host_packed(hj, hr)
    ha, hb, hr3 = env.alloc_host_inputs()
    ha.copy_(hj & 15); hb.copy_(hj >> 4); hr3.copy_(hr)
    for _ in range(iters):
        env.step_host(ha, hb, hr3, narrow=True, zero_copy=True)
    torch.cuda.synchronize()


def k2(kernel, n, K, iters):
    dev = torch.device("cuda", 0)
    env = SoccerVecEnv(n, device=dev, kernel=kernel, rng_mode="philox", seed=0)
    env.reset()
    bufs = (torch.empty((K, n), dtype=torch.int32, device=dev), torch.empty((K, n), dtype=torch.float32, device=dev),
            torch.empty((K, n), dtype=torch.uint8, device=dev))
    for _ in range(iters):
        env.rollout(K, out=bufs)
    torch.cuda.synchronize()


def k1_slip(n, iters):
    dev = torch.device("cuda", 0)
    env = SoccerVecEnv(n, slip_prob=0.2, device=dev, kernel="table", want_reset_obs=False)
    g = torch.Generator(device=dev).manual_seed(0)
    a, b, r = (torch.randint(0, hi, (n,), dtype=torch.uint8, device=dev, generator=g) for hi in (5, 5, 16))
    r32 = torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device=dev, generator=g)
    env.reset(r)
    for _ in range(iters):
        env.step(a, b, r, rng32=r32)
    e2 = SoccerVecEnv(1 << 20, slip_prob=0.2, device=dev, kernel="table", rng_mode="philox")
    e2.reset()
    bufs = (torch.empty((16, 1 << 20), dtype=torch.int32, device=dev), torch.empty((16, 1 << 20), dtype=torch.float32, device=dev),
            torch.empty((16, 1 << 20), dtype=torch.uint8, device=dev))
    for _ in range(iters):
        e2.rollout(16, out=bufs)
    torch.cuda.synchronize()


def k2_slip(n, K, iters):
    dev = torch.device("cuda", 0)
    e2 = SoccerVecEnv(n, slip_prob=0.2, device=dev, kernel="table", rng_mode="philox")
    e2.reset()
    e2.rollout(64, want_streams=False)          # a played-in population
    bufs = (torch.empty((K, n), dtype=torch.int32, device=dev), torch.empty((K, n), dtype=torch.float32, device=dev),
            torch.empty((K, n), dtype=torch.uint8, device=dev))
    for _ in range(iters):
        e2.rollout(K, out=bufs)
    torch.cuda.synchronize()


def replay(kernel, n, T, iters):
    dev = torch.device("cuda", 0)
    env = SoccerVecEnv(n, device=dev, kernel=kernel, want_reset_obs=False)
    a, b, r = (torch.randint(0, hi, (T, n), dtype=torch.uint8, device=dev) for hi in (5, 5, 16))
    out = (torch.empty((T, n), dtype=torch.int32, device=dev), torch.empty((T, n), dtype=torch.float32, device=dev),
           torch.empty((T, n), dtype=torch.uint8, device=dev), None)
    env.reset(r[0])
    for _ in range(iters):
        env.step_many(a, b, r, out=out)
    torch.cuda.synchronize()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="all")
    ap.add_argument("--envs", type=int, default=1 << 24)
    ap.add_argument("--iters", type=int, default=3)
    args = ap.parse_args()
    if args.what in ("k1_table", "all"):
        k1("table", args.envs, args.iters)
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
    if args.what in ("replay", "all"):
        replay("table", 1 << 22, 64, 3)
        replay("table", 4096, 4000, 3)
    print("done")
