"""Group-order experiments with the K1 traffic probe (soccer_bench_stream_mix modes) at several batch sizes."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_soccer_littman94_b200 import _lib
dev = torch.device("cuda", 0)
L = _lib.lib()
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
RING = 4
for n in (1 << 24, 1 << 26):
    g = torch.Generator(device=dev).manual_seed(0)
    state = torch.zeros(n, dtype=torch.int32, device=dev)
    ins = [tuple(torch.randint(0, 255, (n,), dtype=torch.uint8, device=dev, generator=g) for _ in range(3)) for _ in range(RING)]
    outs = [(torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.float32, device=dev),
             torch.empty(n, dtype=torch.uint8, device=dev)) for _ in range(RING)]
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for mode in (0, 1, 2, 0, 1, 2):
        def run(i):
            a, b, r = ins[i % RING]; o, w, f = outs[i % RING]
            _lib.check(L.soccer_bench_stream_mix(C.c_void_p(state.data_ptr()), C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()),
                                                 C.c_void_p(r.data_ptr()), C.c_void_p(o.data_ptr()), C.c_void_p(w.data_ptr()),
                                                 C.c_void_p(f.data_ptr()), n, mode, st), "probe")
        for i in range(5):
            run(i)
        torch.cuda.synchronize(); s0.record()
        for i in range(40):
            run(i)
        s1.record(); torch.cuda.synchronize()
        ms = s0.elapsed_time(s1) / 40
        print(f"n=2^{n.bit_length()-1} mode={mode}: {ms*1e3:.1f} us  {n*20/ms/1e6:.0f} GB/s")
