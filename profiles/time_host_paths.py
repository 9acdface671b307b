"""Closed-loop host-buffer step at 2^24 envs: staged (copy engines, chunked) vs zero copy (the kernel reads / writes
pinned host memory itself), wide and packed streams."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_soccer_littman94_b200.envs import SoccerVecEnv

dev = torch.device("cuda", 0)
for logn in (24, 22, 20, 17):
    n = 1 << logn
    env = SoccerVecEnv(n, device=dev, kernel="table", want_reset_obs=False)
    g = torch.Generator().manual_seed(1)
    arena = os.environ.get("HOST_INPUTS", "arena") == "arena"     # env.alloc_host_inputs() vs tensor.pin_memory()
    ins, pk = [], []
    for _ in range(2):
        vals = [torch.randint(0, hi, (n,), dtype=torch.uint8, generator=g) for hi in (5, 5, 16)]
        if arena:
            bufs, pbufs = env.alloc_host_inputs(), env.alloc_host_inputs(packed=True)
            for b, v in zip(bufs, vals):
                b.copy_(v)
            pbufs[0].copy_(SoccerVecEnv.pack_joint(vals[0], vals[1])); pbufs[1].copy_(vals[2])
        else:
            bufs = tuple(v.pin_memory() for v in vals)
            pbufs = (SoccerVecEnv.pack_joint(vals[0], vals[1]).pin_memory(), vals[2].pin_memory())
        ins.append(bufs); pk.append(pbufs)
    env.reset(ins[0][2].to(dev))

    def run(label, fn, bytes_up, bytes_dn):
        for i in range(2):
            fn(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        R = 8
        for i in range(R):
            fn(i)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / R
        print(f"[{'arena' if arena else 'pin_memory'}] n=2^{logn} {label:28s}: {dt*1e3:7.3f} ms/step  {n/dt/1e9:6.2f} G env-steps/s  up {bytes_up*n/dt/1e9:5.1f} GB/s  down {bytes_dn*n/dt/1e9:5.1f} GB/s", flush=True)

    for ch in (2, 4, 8):
        run(f"staged narrow chunks={ch}", lambda i: env.step_host(*ins[i % 2], narrow=True, n_chunks=ch, zero_copy=False), 3, 4)
    for ch in (1, 2, 4, 8):
        run(f"staged packed chunks={ch}", lambda i: env.step_host_packed(*pk[i % 2], n_chunks=ch, zero_copy=False), 2, 2)
    run("zero-copy narrow", lambda i: env.step_host(*ins[i % 2], narrow=True, zero_copy=True), 3, 4)
    run("zero-copy wide", lambda i: env.step_host(*ins[i % 2], narrow=False, zero_copy=True), 3, 9)
    run("zero-copy packed", lambda i: env.step_host_packed(*pk[i % 2], zero_copy=True), 2, 2)
    for ch in (2, 4, 8, 16, 32):
        run(f"hybrid packed chunks={ch}", lambda i: env.step_host_packed(*pk[i % 2], n_chunks=ch, zero_copy="out"), 2, 2)
    del env
