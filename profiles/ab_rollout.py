"""A/B timing of the K2 kernels for the library given by SOCCER_B200_LIB (BASELINE config 3)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_soccer_littman94_b200.envs import SoccerVecEnv
dev = torch.device("cuda", 0)
for kernel in os.environ.get("AB_KERNELS", "table,rules").split(","):
    for n in (1 << 20, 1 << 21):
        K = 64
        e = SoccerVecEnv(n, device=dev, kernel=kernel, rng_mode="philox", seed=0)
        e.reset()
        bufs = (torch.empty((K, n), dtype=torch.int32, device=dev), torch.empty((K, n), dtype=torch.float32, device=dev),
                torch.empty((K, n), dtype=torch.uint8, device=dev))
        e.rollout(K, out=bufs); torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(8):
            e.rollout(K, out=bufs)
        s1.record(); torch.cuda.synchronize()
        ms = s0.elapsed_time(s1) / 8
        print(os.environ.get("SOCCER_B200_LIB", "default").split("/")[-1], kernel, n, f"{ms:.4f} ms  {n*K/ms/1e6:.1f} G env-steps/s")
