"""K1 with slip_prob > 0 through the shared-memory table: the in-place walk (SOCCER_B200_SLIP_WALK=1) vs the
constant-prefix fast path + deferral queue (slip index), 5x4, device tensors, CUDA events; random states (the env
population after 200 random steps) and the reset population (every env on a start state)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_soccer_littman94_b200.envs import SoccerVecEnv
dev = torch.device("cuda", 0)
for slip in (0.2, 0.5):
    for logn in (22, 24):
        n = 1 << logn
        g = torch.Generator(device=dev).manual_seed(0)
        ins = [tuple(torch.randint(0, hi, (n,), dtype=torch.uint8, device=dev, generator=g) for hi in (5, 5, 16)) +
               (torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device=dev, generator=g),) for _ in range(4)]
        for mode in ("walk", "queued"):
            os.environ["SOCCER_B200_SLIP_WALK"] = "1" if mode == "walk" else "0"
            e = SoccerVecEnv(n, slip_prob=slip, device=dev, kernel="table", want_reset_obs=False)
            e.reset(ins[0][2])
            for i in range(40):                                  # spread the population over the pitch
                a, b, r, r32 = ins[i % 4]
                e.step(a, b, r, rng32=r32)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0.record()
            R = 40
            for i in range(R):
                a, b, r, r32 = ins[i % 4]
                e.step(a, b, r, rng32=r32)
            t1.record()
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / R
            print(f"K1 slip={slip} table {mode:6s} n=2^{logn}: {ms*1e3:.1f} us  {n/ms/1e6:.1f} G env-steps/s  {24*n/ms/1e6:.0f} GB/s (24 B)", flush=True)
            del e
