import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_soccer_littman94_b200.envs import SoccerVecEnv
dev = torch.device("cuda", 0)
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 16
for slip in (0.0, 0.2):
    for n in (1 << 22, (1 << 22) + 1):
        e = SoccerVecEnv(n, width=int(os.environ.get("K2W", "7")), height=int(os.environ.get("K2H", "5")), slip_prob=slip, device=dev, kernel=os.environ.get("K2KERNEL", "rules"), rng_mode="philox")
        e.reset()
        bufs = (torch.empty((K, n), dtype=torch.int32, device=dev), torch.empty((K, n), dtype=torch.float32, device=dev),
                torch.empty((K, n), dtype=torch.uint8, device=dev))
        for _ in range(2): e.rollout(K, out=bufs)
        torch.cuda.synchronize(); s0.record()
        for _ in range(4): e.rollout(K, out=bufs)
        s1.record(); torch.cuda.synchronize()
        ms = s0.elapsed_time(s1) / 4
        print(f"K2 rules slip={slip} n={n} (VEC={'4' if n % 4 == 0 else '1'}): {n*K/ms/1e6:.1f} G env-steps/s")
