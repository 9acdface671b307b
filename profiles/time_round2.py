"""Round-2 kernel timings on one B200 (device-resident tensors, CUDA events, warm-up first):
K1 plain / fused statistics / Philox draws / single-agent (folded policy) / slip 0.2 (injected + Philox),
K2 slip 0 / slip 0.2 (fast path + per-step queue) / in-place walk.   Usage: python profiles/time_round2.py [what ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gym_soccer_littman94_b200.envs import SoccerVecEnv

dev = torch.device("cuda", 0)
PEAK = 6535.4
what = set(sys.argv[1:]) or {"k1", "k2"}


def timed(fn, reps, warm=5):
    for i in range(warm):
        fn(i)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    for i in range(reps):
        fn(i)
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


def report(tag, n_steps, ms, nbytes):
    gbs = n_steps * nbytes / ms / 1e6
    print(f"{tag:58s} {ms * 1e3:9.1f} us  {n_steps / ms / 1e6:8.1f} G env-steps/s  {gbs:7.0f} GB/s ({nbytes} B) = {gbs / PEAK:.3f} of peak",
          flush=True)


if "k1" in what:
    for logn in (24, 26):
        n = 1 << logn
        g = torch.Generator(device=dev).manual_seed(1)
        R = 4
        ins = [tuple(torch.randint(0, hi, (n,), dtype=torch.uint8, device=dev, generator=g) for hi in (5, 5, 16)) for _ in range(R)]
        outs = [(torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.float32, device=dev),
                 torch.empty(n, dtype=torch.uint8, device=dev), None) for _ in range(R)]
        r32 = [torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device=dev, generator=g) for _ in range(2)]
        pol = np.random.RandomState(0).randint(0, 5, 761).astype(np.int8)
        stats = torch.zeros(6, dtype=torch.int64, device=dev)
        cases = [
            ("plain table", dict(kernel="table"), lambda e, i: e.step(*ins[i % R], out=outs[i % R]), 20),
            ("plain table + fused stats", dict(kernel="table"), lambda e, i: e.step(*ins[i % R], out=outs[i % R], stats=stats), 20),
            ("plain rules", dict(kernel="rules"), lambda e, i: e.step(*ins[i % R], out=outs[i % R]), 20),
            ("plain rules + fused stats", dict(kernel="rules"), lambda e, i: e.step(*ins[i % R], out=outs[i % R], stats=stats), 20),
            ("philox table", dict(kernel="table", rng_mode="philox"), lambda e, i: e.step(ins[i % R][0], ins[i % R][1], out=outs[i % R]), 19),
            ("philox table + fused stats", dict(kernel="table", rng_mode="philox"),
             lambda e, i: e.step(ins[i % R][0], ins[i % R][1], out=outs[i % R], stats=stats), 19),
            ("philox rules", dict(kernel="rules", rng_mode="philox"), lambda e, i: e.step(ins[i % R][0], ins[i % R][1], out=outs[i % R]), 19),
            ("single-agent (B folded) table", dict(kernel="table", player_b_policy=pol),
             lambda e, i: e.step(ins[i % R][0], None, ins[i % R][2], out=outs[i % R]), 19),
            ("single-agent (A folded) table philox", dict(kernel="table", player_a_policy=pol, rng_mode="philox"),
             lambda e, i: e.step(None, ins[i % R][1], out=outs[i % R]), 18),
            ("single-agent (B folded) rules/generic", dict(kernel="rules", player_b_policy=pol),
             lambda e, i: e.step(ins[i % R][0], None, ins[i % R][2], out=outs[i % R]), 19),
            ("slip 0.2 table injected rng32", dict(kernel="table", slip_prob=0.2),
             lambda e, i: e.step(*ins[i % R], rng32=r32[i % 2], out=outs[i % R]), 24),
            ("slip 0.2 table philox", dict(kernel="table", slip_prob=0.2, rng_mode="philox"),
             lambda e, i: e.step(ins[i % R][0], ins[i % R][1], out=outs[i % R]), 19),
            ("slip 0.2 table single-agent philox", dict(kernel="table", slip_prob=0.2, rng_mode="philox", player_b_policy=pol),
             lambda e, i: e.step(ins[i % R][0], None, out=outs[i % R]), 18),
        ]
        if logn == 26:
            cases = [c for c in cases if c[0] in ("plain table", "plain table + fused stats", "philox table")]
        for tag, kw, fn, nb in cases:
            e = SoccerVecEnv(n, device=dev, want_reset_obs=False, **kw)
            e.reset(ins[0][2] if e.rng_mode == "injected" else None)
            for i in range(30):                    # play the population in
                fn(e, i)
            ms = timed(lambda i: fn(e, i), 40)
            report(f"K1 {tag} n=2^{logn}", n, ms, nb)
            del e
        del ins, outs, r32
        torch.cuda.empty_cache()

if "k1slip" in what:        # A/B of library variants (SOCCER_B200_LIB): K1 slip 0.2 through the integer fast path
    tag = os.environ.get("SOCCER_B200_LIB", "default").split("/")[-1]
    for logn in (22, 24):
        n = 1 << logn
        g = torch.Generator(device=dev).manual_seed(1)
        ins = [tuple(torch.randint(0, hi, (n,), dtype=torch.uint8, device=dev, generator=g) for hi in (5, 5, 16)) for _ in range(4)]
        outs = [(torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.float32, device=dev),
                 torch.empty(n, dtype=torch.uint8, device=dev), None) for _ in range(4)]
        r32 = [torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device=dev, generator=g) for _ in range(2)]
        for name, kw, fn, nb in (
                ("slip 0.2 injected rng32", dict(slip_prob=0.2), lambda e, i: e.step(*ins[i % 4], rng32=r32[i % 2], out=outs[i % 4]), 24),
                ("slip 0.2 philox", dict(slip_prob=0.2, rng_mode="philox"), lambda e, i: e.step(ins[i % 4][0], ins[i % 4][1], out=outs[i % 4]), 19)):
            e = SoccerVecEnv(n, device=dev, kernel="table", want_reset_obs=False, **kw)
            e.reset(ins[0][2] if e.rng_mode == "injected" else None)
            for i in range(30):
                fn(e, i)
            ms = timed(lambda i: fn(e, i), 40)
            report(f"[{tag}] K1 {name} n=2^{logn}", n, ms, nb)
            del e

if "rules_slip" in what:    # slip_prob = 0.2 on a pitch without a table (7x5): integer thresholds + byte-parallel rules vs the walk
    for tag, envv in (("integer thresholds", {}), ("reference walk", {"SOCCER_B200_SLIP_WALK": "1"})):
        os.environ.update(envv)
        n = 1 << 24
        g = torch.Generator(device=dev).manual_seed(1)
        ins = [tuple(torch.randint(0, hi, (n,), dtype=torch.uint8, device=dev, generator=g) for hi in (5, 5, 16)) for _ in range(2)]
        outs = [(torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.float32, device=dev),
                 torch.empty(n, dtype=torch.uint8, device=dev), None) for _ in range(2)]
        r32 = [torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device=dev, generator=g) for _ in range(2)]
        for name, kw, fn, nb in (
                ("injected rng32", dict(), lambda e, i: e.step(*ins[i % 2], rng32=r32[i % 2], out=outs[i % 2]), 24),
                ("philox", dict(rng_mode="philox"), lambda e, i: e.step(ins[i % 2][0], ins[i % 2][1], out=outs[i % 2]), 19)):
            e = SoccerVecEnv(n, width=7, height=5, slip_prob=0.2, device=dev, kernel="rules", want_reset_obs=False, **kw)
            e.reset(ins[0][2] if e.rng_mode == "injected" else None)
            for i in range(20):
                fn(e, i)
            ms = timed(lambda i: fn(e, i), 10 if envv else 30, warm=2)
            report(f"K1 rules 7x5 slip 0.2 {name} [{tag}] n=2^24", n, ms, nb)
            del e
        del ins, outs, r32
        torch.cuda.empty_cache()
        n, K = 1 << 22, 16
        bufs = (torch.empty((K, n), dtype=torch.int32, device=dev), torch.empty((K, n), dtype=torch.float32, device=dev),
                torch.empty((K, n), dtype=torch.uint8, device=dev))
        e = SoccerVecEnv(n, width=7, height=5, slip_prob=0.2, device=dev, rng_mode="philox", kernel="rules")
        e.reset()
        e.rollout(32, want_streams=False)
        ms = timed(lambda i: e.rollout(K, out=bufs), 4, warm=1)
        report(f"K2 rules 7x5 slip 0.2 [{tag}] n=2^22 K=16", n * K, ms, 9.125)
        del e, bufs
        torch.cuda.empty_cache()
        for k in envv:
            os.environ.pop(k)

if "k2" in what:
    for logn, K in ((20, 64), (21, 64), (22, 16)):
        n = 1 << logn
        bufs = (torch.empty((K, n), dtype=torch.int32, device=dev), torch.empty((K, n), dtype=torch.float32, device=dev),
                torch.empty((K, n), dtype=torch.uint8, device=dev))
        for tag, kw, env in (("slip 0 table", dict(kernel="table"), {}),
                             ("slip 0.2 table integer thresholds", dict(kernel="table", slip_prob=0.2), {}),
                             ("slip 0.2 table in-place walk", dict(kernel="table", slip_prob=0.2), {"SOCCER_B200_SLIP_WALK": "1"}),
                             ("slip 0 rules", dict(kernel="rules"), {}),
                             ("slip 0.2 rules", dict(kernel="rules", slip_prob=0.2), {})):
            if "rules" in tag and logn != 22:
                continue
            os.environ.update(env)
            e = SoccerVecEnv(n, device=dev, rng_mode="philox", **kw)
            e.reset()
            e.rollout(64, want_streams=False)       # play in
            ms = timed(lambda i: e.rollout(K, out=bufs), 6, warm=2)
            report(f"K2 {tag} n=2^{logn} K={K}", n * K, ms, 9.125)
            for k in env:
                os.environ.pop(k)
            del e
        del bufs
        torch.cuda.empty_cache()

if "k2slip" in what:        # A/B of library variants (SOCCER_B200_LIB): K2 slip 0.2, table (5x4) and rules (7x5)
    tag = os.environ.get("SOCCER_B200_LIB", "default").split("/")[-1]
    for logn, K in ((20, 64), (22, 16)):
        n = 1 << logn
        bufs = (torch.empty((K, n), dtype=torch.int32, device=dev), torch.empty((K, n), dtype=torch.float32, device=dev),
                torch.empty((K, n), dtype=torch.uint8, device=dev))
        for name, kw in (("table 5x4", dict(kernel="table")), ("rules 7x5", dict(kernel="rules", width=7, height=5))):
            e = SoccerVecEnv(n, device=dev, rng_mode="philox", slip_prob=0.2, **kw)
            e.reset()
            e.rollout(64, want_streams=False)       # play in
            ms = timed(lambda i: e.rollout(K, out=bufs), 8, warm=2)
            report(f"[{tag}] K2 slip 0.2 {name} n=2^{logn} K={K}", n * K, ms, 9.125)
            del e
        del bufs
        torch.cuda.empty_cache()

if "k2pol" in what:         # K2 with an on-device TABLE policy for player A (the evaluation loop of the planners' policies)
    n, K = 1 << 22, 16
    bufs = (torch.empty((K, n), dtype=torch.int32, device=dev), torch.empty((K, n), dtype=torch.float32, device=dev),
            torch.empty((K, n), dtype=torch.uint8, device=dev))
    for name, kw in (("table 5x4", dict(kernel="table")), ("rules 5x4", dict(kernel="rules")), ("rules 7x5", dict(kernel="rules", width=7, height=5)),
                     ("table 5x4 slip 0.2", dict(kernel="table", slip_prob=0.2)), ("rules 7x5 slip 0.2", dict(kernel="rules", width=7, height=5, slip_prob=0.2))):
        e = SoccerVecEnv(n, device=dev, rng_mode="philox", **kw)
        pol = torch.from_numpy(np.random.RandomState(0).randint(0, 5, e.nS).astype(np.int8)).to(dev)
        e.reset()
        e.rollout(64, want_streams=False, policy_a=pol)
        ms = timed(lambda i: e.rollout(K, out=bufs, policy_a=pol), 8, warm=2)
        report(f"K2 table policy for A, {name} n=2^22 K=16", n * K, ms, 9.125)
        del e
