"""Throughput of the slip_prob > 0 paths (K1 generic with injected 32-bit draws, K1 Philox, K2 rules) and of the
other generic-kernel options, next to the slip-0 fast paths."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_soccer_littman94_b200.envs import SoccerVecEnv
dev = torch.device("cuda", 0)
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s0.record()
    for _ in range(reps):
        fn()
    s1.record()
    torch.cuda.synchronize()
    return s0.elapsed_time(s1) / reps


n = 1 << 22
g = torch.Generator(device=dev).manual_seed(0)
a, b, r = (torch.randint(0, hi, (n,), dtype=torch.uint8, device=dev, generator=g) for hi in (5, 5, 16))
r32 = torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device=dev, generator=g)
for slip in (0.0, 0.2):
    for w, h in ((5, 4), (7, 5)):
        e = SoccerVecEnv(n, width=w, height=h, slip_prob=slip, device=dev, kernel="rules", want_reset_obs=False)
        e.reset(r)
        ms = timeit(lambda: e.step(a, b, r, rng32=r32 if slip else None))
        print(f"K1 injected slip={slip} {w}x{h}: {ms*1e3:.1f} us  {n/ms/1e6:.1f} G env-steps/s")
        if (w, h) == (5, 4) and slip:
            et = SoccerVecEnv(n, width=w, height=h, slip_prob=slip, device=dev, kernel="table", want_reset_obs=False)
            et.reset(r)
            ms = timeit(lambda: et.step(a, b, r, rng32=r32))
            print(f"K1 TABLE    slip={slip} {w}x{h}: {ms*1e3:.1f} us  {n/ms/1e6:.1f} G env-steps/s")
            et = SoccerVecEnv(n, width=w, height=h, slip_prob=slip, device=dev, kernel="table", rng_mode="philox")
            et.reset()
            K = 16
            bufs = (torch.empty((K, n), dtype=torch.int32, device=dev), torch.empty((K, n), dtype=torch.float32, device=dev),
                    torch.empty((K, n), dtype=torch.uint8, device=dev))
            ms = timeit(lambda: et.rollout(K, out=bufs), reps=4)
            print(f"K2 TABLE    slip={slip} {w}x{h}: {ms*1e3:.1f} us  {n*K/ms/1e6:.1f} G env-steps/s")
        e = SoccerVecEnv(n, width=w, height=h, slip_prob=slip, device=dev, kernel="rules", rng_mode="philox", want_reset_obs=False)
        e.reset()
        ms = timeit(lambda: e.step(a, b))
        print(f"K1 philox   slip={slip} {w}x{h}: {ms*1e3:.1f} us  {n/ms/1e6:.1f} G env-steps/s")
        K = 16
        bufs = (torch.empty((K, n), dtype=torch.int32, device=dev), torch.empty((K, n), dtype=torch.float32, device=dev),
                torch.empty((K, n), dtype=torch.uint8, device=dev))
        ms = timeit(lambda: e.rollout(K, out=bufs), reps=4)
        print(f"K2 rules    slip={slip} {w}x{h}: {ms*1e3:.1f} us  {n*K/ms/1e6:.1f} G env-steps/s")
