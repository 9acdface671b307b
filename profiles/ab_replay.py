"""A/B timing of the fused replay (soccer_step_many) for the library given by SOCCER_B200_LIB."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_soccer_littman94_b200.envs import SoccerVecEnv
dev = torch.device("cuda", 0)
tag = os.environ.get("SOCCER_B200_LIB", "default").split("/")[-1]
for n, T in ((4096, 10000), (1 << 14, 4096), (1 << 16, 1024), (1 << 18, 256)):
    a, b, r = (torch.randint(0, hi, (T, n), dtype=torch.uint8, device=dev) for hi in (5, 5, 16))
    out = (torch.empty((T, n), dtype=torch.int32, device=dev), torch.empty((T, n), dtype=torch.float32, device=dev),
           torch.empty((T, n), dtype=torch.uint8, device=dev), None)
    for kernel in ("table", "rules"):
        e = SoccerVecEnv(n, device=dev, kernel=kernel, want_reset_obs=False)
        e.reset(r[0])
        for _ in range(2):
            e.step_many(a, b, r, out=out)
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(8):
            e.step_many(a, b, r, out=out)
        s1.record(); torch.cuda.synchronize()
        ms = s0.elapsed_time(s1) / 8
        print(f"{tag} {kernel} n={n} T={T}: {ms*1e3/T:.4f} us/step  {n*T/ms/1e6:.1f} G env-steps/s  {n*T*12.125/ms/1e6:.0f} GB/s")
