"""Why are small pinned H2D copies slow on this box?  Same copy sizes out of (a) their own pinned allocation,
(b) slices of ONE 512 MB pinned allocation, (c) slices of a 2 MB-aligned MADV_HUGEPAGE mmap registered with
cudaHostRegister.  Also SM zero-copy reads of the same buffers (a trivial torch kernel reading host memory is not
available, so (c) only gets the copy-engine numbers)."""
import ctypes, mmap, torch
dev = torch.device("cuda", 0)
torch.cuda.init()
cudart = torch.cuda.cudart()

def bw(h, d, R):
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(R):
            d.copy_(h, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
    return h.numel() * R / e0.elapsed_time(e1) / 1e6

big = torch.empty(512 << 20, dtype=torch.uint8).pin_memory()
big.fill_(1)
# hugepage mmap
HB = 512 << 20
mm = mmap.mmap(-1, HB + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
try:
    mm.madvise(mmap.MADV_HUGEPAGE)
except Exception as e:
    print("madvise failed", e)
addr = ctypes.addressof(ctypes.c_char.from_buffer(mm))
off = (-addr) % (2 << 20)
import numpy as np
arr = np.frombuffer(mm, dtype=np.uint8, count=HB, offset=off)
arr[:] = 1
rc = cudart.cudaHostRegister(arr.ctypes.data, HB, 0)
print("cudaHostRegister rc", rc)
reg = torch.from_numpy(arr)
print("registered tensor pinned:", reg.is_pinned())
for mb in (1, 2, 8, 16, 32, 64):
    nb = mb << 20
    d = torch.empty(nb, dtype=torch.uint8, device=dev)
    own = torch.empty(nb, dtype=torch.uint8).pin_memory(); own.fill_(1)
    R = max(4, 512 // mb)
    a = bw(own, d, R)
    b = bw(big[64 << 20:(64 << 20) + nb], d, R)
    c = bw(reg[64 << 20:(64 << 20) + nb], d, R)
    b2 = bw(big[(64 << 20) + 4096 * 3:(64 << 20) + 4096 * 3 + nb], d, R)
    print(f"H2D {mb:3d} MB: own alloc {a:5.1f}  slice of 512 MB pinned {b:5.1f} (odd offset {b2:5.1f})  hugepage+register {c:5.1f} GB/s", flush=True)
