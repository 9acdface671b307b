import os, sys, torch
sys.path.insert(0, os.getcwd())
from gym_soccer_littman94_b200.envs import SoccerVecEnv
dev = torch.device("cuda", 0)
n, K = 1 << 22, 16
et = SoccerVecEnv(n, slip_prob=0.2, device=dev, kernel="table", rng_mode="philox")
et.reset()
bufs = (torch.empty((K, n), dtype=torch.int32, device=dev), torch.empty((K, n), dtype=torch.float32, device=dev), torch.empty((K, n), dtype=torch.uint8, device=dev))
for _ in range(3): et.rollout(K, out=bufs)
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); s0.record()
for _ in range(4): et.rollout(K, out=bufs)
s1.record(); torch.cuda.synchronize()
ms = s0.elapsed_time(s1) / 4
print(f"K2 TABLE slip=0.2 n=2^22 K=16: {ms*1e3:.1f} us  {n*K/ms/1e6:.1f} G env-steps/s")
