"""PCIe ceiling of the box: pinned H2D, D2H and both at once, by copy size (torch copies on two streams, CUDA events)."""
import torch
dev = torch.device("cuda", 0)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for mb in (2, 8, 32, 64):
    nb = mb << 20
    h_up, h_dn = torch.empty(nb, dtype=torch.uint8).pin_memory(), torch.empty(nb, dtype=torch.uint8).pin_memory()
    d_up, d_dn = torch.empty(nb, dtype=torch.uint8, device=dev), torch.empty(nb, dtype=torch.uint8, device=dev)
    R = max(4, 512 // mb)
    for mode in ("h2d", "d2h", "both"):
        for rep in range(2):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
            for i in range(R):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        d_up.copy_(h_up, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        h_dn.copy_(d_dn, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{mb:3d} MB x {R:3d} {mode:5s}: {nb * R / ms / 1e6:6.1f} GB/s per direction")
