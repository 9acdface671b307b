"""A/B on B200: K2 on mid-size pitches -- rules inline (soccer_rollout) vs the step table sharded over the shared memory
of a thread-block cluster and read through DSMEM (soccer_rollout_table_cluster).  Checks that both give identical
streams / states first, then times them.   Usage: python profiles/ab_k2_cluster.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gym_soccer_littman94_b200 import _lib
from gym_soccer_littman94_b200.envs import SoccerVecEnv

dev = torch.device("cuda", 0)
L = _lib.lib()
PEAK = 6535.4


def p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


for (w, h) in ((7, 5), (8, 5), (9, 5), (6, 5)):
    pitch = _lib.Pitch(w, h, 0.0)
    nb, cl = C.c_int64(), C.c_int32()
    rc = L.soccer_cluster_table_bytes_host(C.byref(pitch), C.byref(nb), C.byref(cl))
    if rc:
        print(f"{w}x{h}: no cluster table (rc {rc})")
        continue
    table = torch.zeros(nb.value // 2, dtype=torch.int16, device=dev)
    _lib.check(L.soccer_build_cluster_table(C.byref(pitch), p(table), None), "build")
    torch.cuda.synchronize()
    for logn, K in ((20, 64), (22, 16)):
        n = 1 << logn
        seed = 5
        ref = SoccerVecEnv(n, width=w, height=h, device=dev, rng_mode="philox", kernel="rules", seed=seed)
        ref.reset()
        st_idx = torch.empty_like(ref.state)
        _lib.check(L.soccer_convert_state(C.byref(pitch), p(ref.state), p(st_idx), 1, n, None), "convert")
        bufs_r = tuple(torch.empty((K, n), dtype=dt, device=dev) for dt in (torch.int32, torch.float32, torch.uint8))
        bufs_c = tuple(torch.empty((K, n), dtype=dt, device=dev) for dt in (torch.int32, torch.float32, torch.uint8))
        stats_c = torch.zeros(6, dtype=torch.int64, device=dev)
        _, _, _, stats_r = ref.rollout(K, out=bufs_r)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

        def cluster_rollout(step0, stats=None):
            _lib.check(L.soccer_rollout_table_cluster(C.byref(pitch), p(table), p(st_idx), seed, step0, K, 0, p(bufs_c[0]),
                                                      p(bufs_c[1]), p(bufs_c[2]), p(stats), n, st), "cluster rollout")
        cluster_rollout(0, stats_c)
        torch.cuda.synchronize()
        same = all(torch.equal(a, b) for a, b in zip(bufs_r, bufs_c)) and torch.equal(stats_r, stats_c)
        back = torch.empty_like(st_idx)
        _lib.check(L.soccer_convert_state(C.byref(pitch), p(st_idx), p(back), 0, n, None), "convert back")
        same = same and torch.equal(back, ref.state)

        def timed(fn, reps=6):
            for _ in range(2):
                fn()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0.record()
            for _ in range(reps):
                fn()
            t1.record()
            torch.cuda.synchronize()
            return t0.elapsed_time(t1) / reps
        ms_r = timed(lambda: ref.rollout(K, out=bufs_r))
        ms_c = timed(lambda: cluster_rollout(K))
        for tag, ms in (("rules inline", ms_r), (f"cluster table (cluster of {cl.value}, DSMEM)", ms_c)):
            g = n * K / ms / 1e6
            print(f"{w}x{h} n=2^{logn} K={K} {tag:42s} {ms * 1e3:8.1f} us {g:7.1f} G env-steps/s = {g * 9.125 / PEAK:.3f} of HBM peak"
                  f"{'' if same else '   RESULTS DIFFER'}", flush=True)
        del ref, bufs_r, bufs_c
