"""Zero-copy packed step at 2^24 envs: which direction limits it?  joint/draw streams and the result stream placed in
pinned host memory or in HBM independently (soccer_step_table_packed through ctypes)."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_soccer_littman94_b200.envs import SoccerVecEnv
from gym_soccer_littman94_b200 import _lib

dev = torch.device("cuda", 0)
n = 1 << 24
env = SoccerVecEnv(n, device=dev, kernel="table", want_reset_obs=False)
g = torch.Generator().manual_seed(1)
hj, hr = env.alloc_host_inputs(packed=True)
hj.copy_(torch.randint(0, 5, (n,), dtype=torch.uint8, generator=g) | (torch.randint(0, 5, (n,), dtype=torch.uint8, generator=g) << 4))
hr.copy_(torch.randint(0, 16, (n,), dtype=torch.uint8, generator=g))
hres = env._pinned(torch.int16)[0]
dj, dr, dres = hj.to(dev), hr.to(dev), torch.empty(n, dtype=torch.int16, device=dev)
env.reset(dr)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr())
for label, j, r, res in (("in host, out HBM ", hj, hr, dres), ("in HBM,  out host", dj, dr, hres), ("in host, out host", hj, hr, hres),
                         ("in HBM,  out HBM ", dj, dr, dres)):
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        R = 10
        for i in range(R):
            _lib.check(env.lib.soccer_step_table_packed(C.byref(env.pitch), p(env.table), p(env.state), p(j), p(r), p(res), n, st), "x")
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / R
    print(f"{label}: {dt*1e3:7.3f} ms/step  {n/dt/1e9:6.2f} G env-steps/s  {2*n/dt/1e9:5.1f} GB/s per PCIe direction in use")
