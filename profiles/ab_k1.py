"""A/B timing of the K1 kernels for the library given by SOCCER_B200_LIB at several batch sizes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_soccer_littman94_b200.envs import SoccerVecEnv
dev = torch.device("cuda", 0)
tag = os.environ.get("SOCCER_B200_LIB", "default").split("/")[-1]
RING = 4
for kernel in os.environ.get("AB_KERNELS", "table,rules").split(","):
    for n in (1 << 22, 1 << 24, 1 << 26):
        env = SoccerVecEnv(n, device=dev, kernel=kernel, want_reset_obs=False)
        g = torch.Generator(device=dev).manual_seed(0)
        ins = [tuple(torch.randint(0, hi, (n,), dtype=torch.uint8, device=dev, generator=g) for hi in (5, 5, 16)) for _ in range(RING)]
        outs = [(torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.float32, device=dev),
                 torch.empty(n, dtype=torch.uint8, device=dev), None) for _ in range(RING)]
        env.reset(ins[0][2])
        for i in range(5):
            env.step(*ins[i % RING], out=outs[i % RING])
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 50
        s0.record()
        for i in range(K):
            env.step(*ins[i % RING], out=outs[i % RING])
        s1.record(); torch.cuda.synchronize()
        ms = s0.elapsed_time(s1) / K
        print(f"{tag} {kernel} n=2^{n.bit_length()-1}: {ms*1e3:.1f} us  {n/ms/1e6:.1f} G env-steps/s  {n*20/ms/1e6:.0f} GB/s")
        del env, ins, outs
