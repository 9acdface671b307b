"""Store-flavour / slot-order experiments with the K2 write-mix probe, plus torch fill_ and copy_ for reference."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_soccer_littman94_b200 import _lib
dev = torch.device("cuda", 0)
L = _lib.lib()
K = 64
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(fn, reps=16):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s0.record()
    for _ in range(reps):
        fn()
    s1.record()
    torch.cuda.synchronize()
    return s0.elapsed_time(s1) / reps


for n in (1 << 20, 1 << 21):
    state = torch.zeros(n, dtype=torch.int32, device=dev)
    obs = torch.empty((K, n), dtype=torch.int32, device=dev)
    rew = torch.empty((K, n), dtype=torch.float32, device=dev)
    flg = torch.empty((K, n), dtype=torch.uint8, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for mode in range(6):
        ms = timeit(lambda: _lib.check(L.soccer_bench_rollout_probe(
            C.c_void_p(state.data_ptr()), K, C.c_void_p(obs.data_ptr()), C.c_void_p(rew.data_ptr()),
            C.c_void_p(flg.data_ptr()), n, mode, st), "probe"))
        print(f"n={n} mode={mode} {ms:.4f} ms {n*K*9.125/ms/1e6:.0f} GB/s")
    ms = timeit(lambda: (obs.fill_(1), rew.fill_(1.0), flg.fill_(1)))
    print(f"n={n} torch fill_ x3 {ms:.4f} ms {n*K*9/ms/1e6:.0f} GB/s")
    ms = timeit(lambda: rew.copy_(obs))
    print(f"n={n} torch copy_ i32->f32 {ms:.4f} ms {n*K*8/ms/1e6:.0f} GB/s")
big = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
ms = timeit(lambda: big.fill_(3))
print(f"fill_ 1 GiB {ms:.4f} ms {(1<<30)/ms/1e6:.0f} GB/s")
big2 = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
ms = timeit(lambda: big2.copy_(big))
print(f"copy_ 1 GiB {ms:.4f} ms {2*(1<<30)/ms/1e6:.0f} GB/s")
