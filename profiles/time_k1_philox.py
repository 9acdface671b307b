"""K1 with on-device Philox draws: rules kernel (soccer_step_philox) vs shared-memory table
(soccer_step_table_philox), 5x4 slip 0, device-resident tensors, CUDA events.  Usage: python profiles/time_k1_philox.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_soccer_littman94_b200.envs import SoccerVecEnv

dev = torch.device("cuda", 0)
for logn in (22, 24):
    n = 1 << logn
    g = torch.Generator(device=dev).manual_seed(1)
    ins = [tuple(torch.randint(0, 5, (n,), dtype=torch.uint8, device=dev, generator=g) for _ in range(2)) for _ in range(4)]
    for kernel in ("rules", "table"):
        e = SoccerVecEnv(n, device=dev, kernel=kernel, rng_mode="philox", want_reset_obs=False)
        e.reset()
        for i in range(5):
            e.step(*ins[i % 4])
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record()
        R = 40
        for i in range(R):
            e.step(*ins[i % 4])
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / R
        print(f"K1 philox {kernel} n=2^{logn}: {ms*1e3:.1f} us  {n/ms/1e6:.1f} G env-steps/s  {19*n/ms/1e6:.0f} GB/s (19 B)")
        del e
