#!/usr/bin/env python
"""bench.py -- lock-step env-steps/s of the Littman'94 soccer step path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...     (N > 1)

Workload (config.workload): the K1 streaming step -- host-supplied joint actions, injected 2-bit
draws, auto-reset fused -- over 2^24 lock-step 5x4 envs PER GPU, the size BASELINE.md grades the
HBM roofline on (BASELINE.json configs[1] semantics; configs[1]'s own N = 4096 moves 82 KB per
step and is launch-latency bound, it is reported under "extra").  One "step" = one kernel launch
advancing every env by one transition.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "lockstep_env_steps_per_sec"
UNIT = "env-steps/s"
# all-reduced statistics of 2^24 GLOBAL envs (5x4, slip 0, Philox seed 20261018, reset() + 64 lock-steps of uniform play):
# computed by the CPU oracle (oracle/make_shard_invariant.py) -- Philox is keyed by the global env id, so every sharding of
# that batch over 1, 2, 4 or 8 GPUs must reproduce this vector exactly
SHARD_INVARIANT = {"envs": 1 << 24, "K": 64, "seed": 20261018,
                   "stats": [29943569, 14967803, 14975766, 0, 1073741824, 590049320]}
BYTES_PER_ENV_STEP = 20   # state u32 r+w (8) + act_a, act_b, rng u8 (3) + obs i32 (4) + reward f32 (4) + flags u8 (1)


_JSON_OUT = sys.stdout


def emit_json(line):
    _JSON_OUT.write(line + "\n")
    _JSON_OUT.flush()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel_prefix, exact=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
    committed ncu --set full capture (profiles/ncu_traffic.json); None if there is none.  `exact`: the instantiation
    the timed region launches (preferred); else the first instantiation of `kernel_prefix`."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            recs = {name.split("::")[-1]: rec for name, rec in json.load(f).items()}
        if exact in recs:
            return recs[exact]["dram_bytes_per_launch"], recs[exact]["source"] + " [" + exact + "]"
        for name, rec in recs.items():
            if name.startswith(kernel_prefix + "<"):
                return rec["dram_bytes_per_launch"], rec["source"] + " [" + name + "]"
    except Exception:  # noqa: BLE001
        pass
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons, pw = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1]); pw.append(float(r[2]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "sm_mhz_min": sm[0] if sm else None, "power_w_max": max(pw) if pw else None}


def time_cpu_port(total_budget_s, n_threads, T=64, N=1 << 15):
    """The oracle port (oracle/soccer_oracle.c: table build once, then lookup + categorical draw
    per step, like the reference) timed on host cores.  Returns (env_steps_per_s, sample text)."""
    import numpy as np
    from oracle import soccer_oracle as so
    m = so.OracleModel(5, 4, 0.0)
    rs = np.random.RandomState(123)
    act_a = rs.randint(0, 5, (T, N)).astype(np.uint8)
    act_b = rs.randint(0, 5, (T, N)).astype(np.uint8)
    rng8 = rs.randint(0, 16, (T, N)).astype(np.uint8)
    states = np.zeros(N, so.STATE_DTYPE)
    for i in range(N):
        states[i] = m.isd[i & 3][1]
    ts = np.zeros(N, np.int32)
    out = (np.zeros((T, N), np.int32), np.zeros((T, N), np.float32), np.zeros((T, N), np.uint8), None)   # pre-faulted
    m.rollout_injected(states, ts, act_a, act_b, rng8, n_threads=n_threads, out=out)   # warm
    reps, t0 = 0, time.perf_counter()
    while True:
        m.rollout_injected(states, ts, act_a, act_b, rng8, n_threads=n_threads, out=out)
        reps += 1
        el = time.perf_counter() - t0
        if el >= total_budget_s or reps >= 4096:
            break
    return reps * T * N / el, f"{reps} x ({N} envs x {T} lock-steps), {el:.1f} s, 5x4 slip 0, injected draws"


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, timed on this box's host
    cores.  The reference itself is Python and cannot travel to the GPU box, so this times the C
    oracle port of it (kind = "port") with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    T, N = 64, 1 << 18          # 16.8 M env-steps per bench step: ~50 ms on 16 cores, a stable sample
    import numpy as np
    from oracle import soccer_oracle as so
    m = so.OracleModel(5, 4, 0.0)
    rs = np.random.RandomState(123)
    act_a = rs.randint(0, 5, (T, N)).astype(np.uint8)
    act_b = rs.randint(0, 5, (T, N)).astype(np.uint8)
    rng8 = rs.randint(0, 16, (T, N)).astype(np.uint8)
    states = np.zeros(N, so.STATE_DTYPE)
    for i in range(N):
        states[i] = m.isd[i & 3][1]
    ts = np.zeros(N, np.int32)
    out = (np.zeros((T, N), np.int32), np.zeros((T, N), np.float32), np.zeros((T, N), np.uint8), None)   # pre-faulted
    for _ in range(max(args.warmup, 1)):
        m.rollout_injected(states, ts, act_a, act_b, rng8, n_threads=cores, out=out)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m.rollout_injected(states, ts, act_a, act_b, rng8, n_threads=cores, out=out)
    el = time.perf_counter() - t0
    v = args.steps * T * N / el
    sample = f"each step = {N} envs x {T} lock-steps through the oracle port, {cores} threads"
    emit_json(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/u32 integer state + fp64 categorical draw", "data": "synthetic",
        "config": {"workload": "K1 lock-step step, 5x4 pitch, slip 0, host-supplied joint actions, injected draws, "
                               "auto-reset; CPU sample of the 2^24-env workload", "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def host_info():
    model = "?"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    import numpy
    return {"cpu": model, "logical_cpus": os.cpu_count(), "python": sys.version.split()[0], "numpy": numpy.__version__}


def bind_near_gpu(local):
    """Pin this rank to the CPU cores NVML reports as local to its GPU, BEFORE any pinned host buffer is allocated
    (first touch puts the buffers on that NUMA node): the end-to-end path is PCIe- and host-memory-bound, and with
    several ranks per box a buffer on the far socket costs a trip over the inter-socket link.  Best effort."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:  # noqa: BLE001
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {i * 64 + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:  # noqa: BLE001
        pass
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist

    from gym_soccer_littman94_b200.envs import SoccerVecEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = bind_near_gpu(local) if (world > 1 and not args.no_bind) else None
    p2p = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # the one collective of the path as a hand-written kernel over NVLink peer memory (falls back to NCCL where the
        # symmetric-memory rendezvous is not available); --collective nccl forces the library all-reduce
        from gym_soccer_littman94_b200.dist import P2PStatsAllReduce
        if args.collective == "p2p":
            p2p = P2PStatsAllReduce.create(dev)
    from gym_soccer_littman94_b200.dist import allreduce_stats
    collective = "none (1 GPU)" if world == 1 else ("soccer_stats_allreduce_p2p (peer-memory kernel over NVLink)" if p2p is not None
                                                    else "ncclAllReduce (torch.distributed)")
    N = args.envs_per_gpu
    K, W = args.steps, max(args.warmup, 3)
    RING = 4

    env = SoccerVecEnv(N, device=dev, kernel=args.kernel, rng_mode="injected", env_id_base=rank * N,
                       want_reset_obs=False)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)

    def rnd(hi):
        return torch.randint(0, hi, (N,), dtype=torch.uint8, device=dev, generator=g)
    ins = [(rnd(5), rnd(5), rnd(16)) for _ in range(RING)]
    outs = [(torch.empty(N, dtype=torch.int32, device=dev), torch.empty(N, dtype=torch.float32, device=dev),
             torch.empty(N, dtype=torch.uint8, device=dev), None) for _ in range(RING)]
    env.reset(rnd(16))
    # play the population in: after reset() every env sits on one of FOUR start observations (4 table rows for 2^24 envs),
    # which is not the workload -- 128 untimed steps spread it over the state space (mean episode length 34)
    for i in range(128):
        env.step(*ins[i % RING], out=outs[i % RING])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (value) + per-launch roofline
    # Episode statistics are accumulated INSIDE the K1 kernel (soccer_step_args.stats: byte-parallel flag counts, one
    # warp reduction + 6 atomics per CTA at the end) at every N, so a step costs the same on 1 GPU and on 8; at N > 1
    # the one collective of the path -- the sum all-reduce of the 48-byte vector -- follows the last K1 directly.
    stats = torch.zeros(6, dtype=torch.int64, device=dev)
    for i in range(W):
        env.step(*ins[i % RING], out=outs[i % RING], stats=stats)
    if world > 1:
        barrier()                                      # the peer-memory kernel has a watchdog: enter it together
        allreduce_stats(stats.clone(), p2p)            # warm the collective
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    # (1) the timed region: EXACTLY K steps between two events, nothing else in the stream
    start, k_done, end = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    barrier()
    # the W warm-up steps run DIRECTLY ahead of the timed ones, in the same stream: the clock sampler's start-up above
    # left the GPU idle for 0.3 s, and the first launches after an idle gap run ~10 % slower than steady state
    for i in range(W):
        env.step(*ins[i % RING], out=outs[i % RING], stats=stats)
    stats.zero_()
    if world > 1:
        # device-side rendezvous right before the start event: the host-side barrier lets the ranks go tens of
        # microseconds apart, which the closing all-reduce would otherwise charge to the early ranks' K steps
        dist.all_reduce(torch.zeros(1, device=dev))
    start.record()
    for i in range(K):
        env.step(*ins[i % RING], out=outs[i % RING], stats=stats)
    k_done.record()
    if world > 1:
        allreduce_stats(stats, p2p)                    # the one collective of the path, same stream
    end.record()
    barrier()
    total_ms = start.elapsed_time(end)
    kern_ms = start.elapsed_time(k_done) / K          # mean launch duration of the dominant kernel
    tail_us = k_done.elapsed_time(end) * 1e3          # what follows the last K1 inside the timed region (the all-reduce)
    stats_timed = [int(x) for x in stats.cpu()]       # episodes, goals_A, goals_B, truncations, env-steps, sum of lengths
    # (2) a second pass with an event after every launch, only for the per-launch spread
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    ev[0].record()
    for i in range(K):
        env.step(*ins[i % RING], out=outs[i % RING], stats=stats)
        ev[i + 1].record()
    torch.cuda.synchronize()
    per_launch = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
    # (3) the same launches back to back for about a second: the clock / throttle samples below are taken under THIS
    # load (the K-step region above lasts a few milliseconds, less than one nvidia-smi sampling period)
    n_sus = max(K, int(1000.0 / max(kern_ms, 1e-3)))
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record()
    for i in range(n_sus):
        env.step(*ins[i % RING], out=outs[i % RING], stats=stats)
    u1.record()
    torch.cuda.synchronize()
    sustained = {"steps": n_sus, "seconds": u0.elapsed_time(u1) * 1e-3,
                 "env_steps_per_s_per_gpu": N * n_sus / (u0.elapsed_time(u1) * 1e-3)}
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["sampled_over"] = "the timed K steps, the per-launch pass and %d further back-to-back launches (%.2f s)" % (
            n_sus, sustained["seconds"])
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    tails = [tail_us]
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # every rank's tail: the rank that arrives last waits for nobody (its tail = the all-reduce alone), the others
        # also wait for it (rank skew over the K steps)
        tt = torch.zeros(world, dtype=torch.float64, device=dev)
        tt[rank] = tail_us
        dist.all_reduce(tt)
        tails = [float(x) for x in tt.cpu()]
    total_ms = float(t.item())
    value = world * N * K / (total_ms * 1e-3)
    peak, peak_src = measured_peaks()
    achieved = BYTES_PER_ENV_STEP * N / (kern_ms * 1e-3) / 1e9

    # ---------------- end to end: pinned host buffers in, pinned host buffers out, every step
    # ---------------- the practical HBM ceiling for K1's traffic mix (7 B read / 13 B written per env):
    # the same streams, access pattern and launch shape with the game logic removed
    import ctypes as C
    from gym_soccer_littman94_b200 import _lib
    scratch_state = env.state.clone()
    cur_stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def probe(i):
        a_, b_, r_ = ins[i % RING]
        o_, w_, f_, _ = outs[i % RING]
        _lib.check(env.lib.soccer_bench_stream_mix(
            C.c_void_p(scratch_state.data_ptr()), C.c_void_p(a_.data_ptr()), C.c_void_p(b_.data_ptr()),
            C.c_void_p(r_.data_ptr()), C.c_void_p(o_.data_ptr()), C.c_void_p(w_.data_ptr()), C.c_void_p(f_.data_ptr()),
            N, 0, cur_stream), "soccer_bench_stream_mix")
    for i in range(W):
        probe(i)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(K):
        probe(i)
    p1.record()
    barrier()
    probe_gbs = BYTES_PER_ENV_STEP * N / (p0.elapsed_time(p1) / K * 1e-3) / 1e9
    del scratch_state

    # every step is a closed loop: upload this step's actions/draws, step, download obs/reward/flags,
    # and wait for them (a host-side policy needs them to pick the next actions)
    h_in = []
    for i in range(2):
        bufs = env.alloc_host_inputs()              # pinned, huge-page backed (soccer_host_alloc)
        for hbuf, x in zip(bufs, ins[i]):
            hbuf.copy_(x)
        h_in.append(bufs)
    Ke = max(3, min(K, 10))
    e2e = {}
    for narrow in (True, False):
        for i in range(2):
            env.step_host(*h_in[i % 2], narrow=narrow, n_chunks=args.chunks)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream(dev)
        e0.record()
        for i in range(Ke):
            h_obs, h_rew, h_flg = env.step_host(*h_in[i % 2], narrow=narrow, n_chunks=args.chunks)   # syncs
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e[narrow] = (world * N * Ke / (float(t.item()) * 1e-3), int(h_obs[:1024].to(torch.int64).sum()))
    e2e_value, checksum = e2e[True]      # the result really is on the host
    # the same closed loop with the packed streams (joint-action byte + draw byte up, one 16-bit result word down)
    e2e_packed = None
    if env.kernel == "table":
        h_pk = []
        for i in range(2):
            bufs = env.alloc_host_inputs(packed=True)
            bufs[0].copy_(SoccerVecEnv.pack_joint(h_in[i][0], h_in[i][1])); bufs[1].copy_(h_in[i][2])
            h_pk.append(bufs)
        for i in range(2):
            env.step_host_packed(*h_pk[i % 2], n_chunks=args.chunks)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(Ke):
            h_res = env.step_host_packed(*h_pk[i % 2], n_chunks=args.chunks)   # syncs
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_packed = {"value": world * N * Ke / (float(t.item()) * 1e-3), "h2d_bytes_per_step": 2 * N,
                      "d2h_bytes_per_step": 2 * N, "checksum": int((h_res[:1024].to(torch.int64) & 0xFFF).sum()),
                      "api": "SoccerVecEnv.step_host_packed: joint-action byte (aa | ab << 4) + draw byte up, one int16 "
                             "result word (obs | terminated << 12 | truncated << 13 | reward << 14) down"}
        # -- verified: one more end-to-end step whose host-side result words are compared, on a random sample of envs,
        #    with the CPU oracle stepping the very states (the oracle is the checker here, never the path measured)
        try:
            import numpy as np
            from oracle import soccer_oracle as so
            m = so.OracleModel(5, 4, 0.0)
            tab = np.zeros(m.nS, so.STATE_DTYPE)
            for o in range(1, m.nS):
                tab[o] = m.obs_to_state(o)
            idx = np.sort(np.random.RandomState(99 + rank).choice(N, 1 << 16, replace=False))
            tidx = torch.from_numpy(idx).to(dev)
            st_before = env.state[tidx].cpu().numpy()
            jb, rb = h_pk[0]
            h_res = env.step_host_packed(jb, rb, n_chunks=args.chunks)
            got = h_res[torch.from_numpy(idx)].numpy().astype(np.int32)
            jn, rn = jb.numpy()[idx], rb.numpy()[idx]
            states, ts = tab[st_before & 0xFFFF].copy(), ((st_before >> 16) & 0xFF).astype(np.int32)
            eo, er, ef, _ = m.rollout_injected(states, ts, (jn & 15)[None, :].copy(), (jn >> 4)[None, :].copy(), rn[None, :].copy(),
                                               want_reset_obs=False)
            ok = (np.array_equal(got & 0xFFF, eo[0]) and np.array_equal((got >> 14), er[0].astype(np.int32))
                  and np.array_equal((got >> 12) & 3, ef[0].astype(np.int32)))
            e2e_packed["verified"] = bool(ok)
            e2e_packed["verified_how"] = ("one further step_host_packed call: result words of 65,536 randomly chosen envs == the "
                                          "CPU oracle stepping the same states with the same actions and draws")
        except Exception as e:  # noqa: BLE001
            e2e_packed["verified"] = False
            e2e_packed["verified_error"] = repr(e)
        # -- the link's own rate for exactly these byte counts: cudaMemcpyAsync H2D of 2N bytes and D2H of 2N bytes, alone
        #    and both directions at once (two streams), from / to the same huge-page-backed pinned arena; then the same
        #    duplex copy on ALL ranks at once (the host side of the box shared by N GPUs)
        link = {}
        try:
            d_up = torch.empty(2 * N, dtype=torch.uint8, device=dev)
            d_dn = torch.empty(2 * N, dtype=torch.uint8, device=dev)
            (h_up,) = env._pinned(torch.int16)
            (h_dn,) = env._pinned(torch.int16)
            h_up8, h_dn8 = h_up.view(torch.uint8), h_dn.view(torch.uint8)
            s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

            def copy_rate(up, dn, reps=6):
                torch.cuda.synchronize()
                c0, c1, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                c0.record()
                s_up.wait_event(c0); s_dn.wait_event(c0)
                for _ in range(reps):
                    if up:
                        with torch.cuda.stream(s_up):
                            d_up.copy_(h_up8, non_blocking=True)
                    if dn:
                        with torch.cuda.stream(s_dn):
                            h_dn8.copy_(d_dn, non_blocking=True)
                with torch.cuda.stream(s_up):
                    c1.record()
                with torch.cuda.stream(s_dn):
                    c2.record()
                torch.cuda.synchronize()
                ms = max(c0.elapsed_time(c1), c0.elapsed_time(c2)) / reps
                return 2 * N / (ms * 1e-3) / 1e9            # GB/s per direction
            copy_rate(True, True, 2)
            link["h2d_alone_gbs"] = copy_rate(True, False)
            link["d2h_alone_gbs"] = copy_rate(False, True)
            link["duplex_gbs_per_direction"] = copy_rate(True, True)
            barrier()
            allr = copy_rate(True, True)
            barrier()
            t = torch.tensor([allr], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
            link["duplex_gbs_per_direction_all_ranks_at_once_min"] = float(t.item())
            del d_up, d_dn, h_up, h_dn
        except Exception as e:  # noqa: BLE001
            link["error"] = repr(e)
        per_rank_gbs = e2e_packed["value"] / world * 2 / 1e9          # bytes per direction per second on one rank's link
        e2e_packed["roofline"] = dict(
            link, bound="pcie", achieved_gbs_per_direction=per_rank_gbs,
            frac_of_duplex_copy_rate=(per_rank_gbs / link["duplex_gbs_per_direction"]) if link.get("duplex_gbs_per_direction") else None,
            frac_of_duplex_copy_rate_all_ranks=(per_rank_gbs / link["duplex_gbs_per_direction_all_ranks_at_once_min"])
            if link.get("duplex_gbs_per_direction_all_ranks_at_once_min") else None,
            note="closed loop: every step waits for its results, so one launch + one synchronize (~10 us) sit on top of the "
                 "transfer; the kernel reads / writes the pinned host buffers itself (zero copy), both directions at once")
        # -- the same closed loop in Philox mode: no draw stream, 1 byte up + 2 bytes down per env (fewer host bytes for
        #    boxes whose host side, not the link, is the limit at N > 2)
        try:
            eph = SoccerVecEnv(N, device=dev, kernel="table", rng_mode="philox", env_id_base=rank * N, seed=1, want_reset_obs=False)
            eph.reset()
            for i in range(2):
                eph.step_host_packed(h_pk[i % 2][0])
            barrier()
            e0.record()
            for i in range(Ke):
                eph.step_host_packed(h_pk[i % 2][0])
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_packed["philox"] = {"value": world * N * Ke / (float(t.item()) * 1e-3), "h2d_bytes_per_step": N,
                                    "d2h_bytes_per_step": 2 * N,
                                    "api": "step_host_packed with rng_mode='philox': joint-action byte up, int16 result word down"}
            del eph
        except Exception as e:  # noqa: BLE001
            e2e_packed["philox"] = {"error": repr(e)}
        del h_pk
    # -- the registration default slip_prob = 0.2 (gym_soccer/__init__.py:8) from host buffers: natural dtypes, zero copy
    e2e_slip = None
    try:
        n_s = min(N, 1 << 22)
        es = SoccerVecEnv(n_s, slip_prob=0.2, device=dev, kernel="auto", want_reset_obs=False)
        sa, sb, sr = es.alloc_host_inputs()
        (s32,) = es._pinned(torch.int32)
        sa.copy_(h_in[0][0][:n_s]); sb.copy_(h_in[0][1][:n_s]); sr.copy_(h_in[0][2][:n_s])
        s32.copy_(torch.randint(-2**31, 2**31 - 1, (n_s,), dtype=torch.int32))
        es.reset(sr.to(dev))
        for i in range(2):
            es.step_host(sa, sb, sr, rng32=s32)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(Ke):
            es.step_host(sa, sb, sr, rng32=s32)
        el = time.perf_counter() - t0
        e2e_slip = {"value_per_gpu": n_s * Ke / el, "envs": n_s, "h2d_bytes_per_step": 7 * n_s, "d2h_bytes_per_step": 9 * n_s,
                    "api": "step_host(act_a, act_b, rng8, rng32=...) on a slip_prob = 0.2 env: int32 obs / float32 reward / "
                           "uint8 flags down"}
        del es
    except Exception as e:  # noqa: BLE001
        e2e_slip = {"error": repr(e)}

    # ---------------- BASELINE configs 3 / 4 on EVERY rank: fused K = 64 rollouts (2^20 and 2^21 envs per GPU,
    # global env ids rank * n + local), 16 launches back to back = 1024 steps, then ONE all-reduce of the
    # 48-byte statistics vector (inside the timed region); plus the write-only traffic probe of the same shape
    rollouts = {}
    try:
        del ins, outs
        torch.cuda.empty_cache()
        K3, L3 = 64, 16
        for tag, n3 in (("config3_rollout_2^20_K64", 1 << 20), ("config4_rollout_2^21_per_gpu_K64", 1 << 21)):
            e3 = SoccerVecEnv(n3, device=dev, kernel=args.kernel, rng_mode="philox", seed=0, env_id_base=rank * n3)
            e3.reset()
            bufs = (torch.empty((K3, n3), dtype=torch.int32, device=dev), torch.empty((K3, n3), dtype=torch.float32, device=dev),
                    torch.empty((K3, n3), dtype=torch.uint8, device=dev))
            for _ in range(3):
                e3.rollout(K3, out=bufs)
            st3 = torch.zeros(6, dtype=torch.int64, device=dev)
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(L3):
                e3.rollout(K3, out=bufs, stats=st3)
            if world > 1:
                allreduce_stats(st3, p2p)
            s1.record()
            barrier()
            t = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item()) / L3
            ar_us = None
            if world > 1:                       # the one collective of the path, timed on its own (10 back to back)
                ar_us = {}
                for name, fn in (("nccl", lambda x: dist.all_reduce(x)),) + ((("p2p_kernel", lambda x: p2p(x)),) if p2p else ()):
                    tmp = st3.clone()
                    fn(tmp)
                    barrier()
                    s0.record()
                    for _ in range(10):
                        fn(tmp)
                    s1.record()
                    barrier()
                    ar_us[name] = s0.elapsed_time(s1) * 100.0
            scratch = e3.state.clone()

            def rprobe():
                _lib.check(env.lib.soccer_bench_rollout_probe(
                    C.c_void_p(scratch.data_ptr()), K3, C.c_void_p(bufs[0].data_ptr()), C.c_void_p(bufs[1].data_ptr()),
                    C.c_void_p(bufs[2].data_ptr()), n3, 0, cur_stream), "soccer_bench_rollout_probe")
            for _ in range(3):
                rprobe()
            barrier()
            s0.record()
            for _ in range(L3):
                rprobe()
            s1.record()
            barrier()
            pms = s0.elapsed_time(s1) / L3
            v3 = world * n3 * K3 / (ms * 1e-3)
            gbs = n3 * K3 * 9.125 / (ms * 1e-3) / 1e9
            pgbs = n3 * K3 * 9.125 / (pms * 1e-3) / 1e9
            rollouts[tag] = {"kernel": e3.kernel, "launches": L3, "ms_per_launch": ms, "env_steps_per_s": v3,
                             "per_gpu_hbm_gbs_at_9.125B": gbs, "frac_of_hbm_peak": gbs / peak,
                             "write_mix_probe_gbs": pgbs, "kernel_over_probe": gbs / pgbs,
                             "stats_allreduce": [int(x) for x in st3.cpu()], "allreduce_us": ar_us,
                             "note": "max over ranks; the all-reduce of the statistics vector is inside the timed region"}
            del e3, bufs, scratch
    except Exception as e:  # noqa: BLE001
        rollouts["error"] = repr(e)

    # ---------------- hardware proof that trajectories do not depend on the GPU count: the SAME 2^24 global envs, sharded
    # over the `world` GPUs of this run (contiguous global ids, env_id_base = rank * n), K2 in Philox mode, one NCCL
    # all-reduce of the statistics; the result must be the constant the CPU oracle computed for the unsharded batch
    shard = {"expected": SHARD_INVARIANT["stats"], "global_envs": SHARD_INVARIANT["envs"], "K": SHARD_INVARIANT["K"],
             "seed": SHARD_INVARIANT["seed"], "shards": world}
    try:
        n_sh = SHARD_INVARIANT["envs"] // world
        esh = SoccerVecEnv(n_sh, device=dev, kernel=args.kernel, rng_mode="philox", seed=SHARD_INVARIANT["seed"],
                           env_id_base=rank * n_sh)
        esh.reset()
        _, _, _, st_sh = esh.rollout(SHARD_INVARIANT["K"], want_streams=False)
        if world > 1:
            barrier()
            allreduce_stats(st_sh, p2p)
        shard["collective"] = collective
        shard["stats_allreduce"] = [int(x) for x in st_sh.cpu()]
        shard["shard_invariant"] = shard["stats_allreduce"] == SHARD_INVARIANT["stats"]
        del esh
    except Exception as e:  # noqa: BLE001
        shard["error"] = repr(e)
        shard["shard_invariant"] = False

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- extra: BASELINE config 2 (reported, not the headline)
    extra = dict(rollouts)
    try:
        # BASELINE config 1 through the single-env drop-in (the reference's own loop): every
        # step() is one kernel launch on a 1-env batch + a stream synchronize, the draws come from
        # np.random.RandomState like the reference's
        import numpy as np
        from gym_soccer_littman94_b200.envs import SoccerSimultaneousEnv
        c1 = {}
        for mode in ("speculative", "speculative_slip_0.2", "zero_copy", "staged"):
            # speculative (the default): one launch per step steps the new state for all 25 joint actions x 4 draws
            # while the Python loop is busy (soccer_step_speculate); zero_copy / staged: launch and wait per step
            os.environ["SOCCER_B200_SINGLE_ENV_STAGED"] = "1" if mode == "staged" else "0"
            os.environ["SOCCER_B200_SINGLE_ENV_SPECULATE"] = "1" if mode.startswith("speculative") else "0"
            e1 = SoccerSimultaneousEnv(5, 4, slip_prob=0.2 if mode.endswith("0.2") else 0.0, seed=0, device=dev)
            acts = np.random.RandomState(123).randint(0, 5, (20000, 2))
            e1.reset()
            t0 = time.perf_counter()
            n_ep = 0
            for aa, ab in acts:
                _, _, dn, tr, _ = e1.step({'player_a': int(aa), 'player_b': int(ab)})
                if dn['player_a'] or tr['player_a']:
                    e1.reset()
                    n_ep += 1
            c1[mode] = {"steps_per_s": len(acts) / (time.perf_counter() - t0), "episodes": n_ep}
        os.environ.pop("SOCCER_B200_SINGLE_ENV_STAGED", None)
        os.environ.pop("SOCCER_B200_SINGLE_ENV_SPECULATE", None)
        extra["config1_single_env_dropin"] = dict(c1, note="20,000 step() calls of ONE env through the reference's class "
                                                  "surface; latency-bound: speculative = one launch per step that steps the "
                                                  "new state for all 25 joint actions x 4 draws ahead of the next call (slip 0.2: the 25 joint actions with the "
                                                  "draw the env's generator is going to make), "
                                                  "zero_copy / staged = one launch + one wait per step; reported next to "
                                                  "cpu_baseline.python_reference (the unmodified reference's loop on this host)")
        # K1 with a state tensor far beyond the L2 (2^26 envs: 268 MB of state, 1.34 GB per step): the DRAM-only figure
        n6 = 1 << 26
        e6 = SoccerVecEnv(n6, device=dev, kernel=args.kernel, want_reset_obs=False)
        g6 = torch.Generator(device=dev).manual_seed(7)
        in6 = [tuple(torch.randint(0, hi, (n6,), dtype=torch.uint8, device=dev, generator=g6) for hi in (5, 5, 16)) for _ in range(2)]
        out6 = [(torch.empty(n6, dtype=torch.int32, device=dev), torch.empty(n6, dtype=torch.float32, device=dev),
                 torch.empty(n6, dtype=torch.uint8, device=dev), None) for _ in range(2)]
        e6.reset(in6[0][2])
        for i in range(3):
            e6.step(*in6[i % 2], out=out6[i % 2])
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        q0.record()
        for i in range(20):
            e6.step(*in6[i % 2], out=out6[i % 2])
        q1.record()
        torch.cuda.synchronize()
        ms6 = q0.elapsed_time(q1) / 20
        extra["k1_2^26_envs"] = {"kernel": e6.kernel, "env_steps_per_s": n6 / (ms6 * 1e-3),
                                 "hbm_gbs_at_20B": n6 * 20 / (ms6 * 1e-3) / 1e9, "frac_of_hbm_peak": n6 * 20 / (ms6 * 1e-3) / 1e9 / peak}
        del e6, in6, out6
        torch.cuda.empty_cache()
        # K1 variants at 2^24 envs: slip_prob = 0.2 with injected 32-bit step draws (24 B per
        # env-step) and slip 0 with on-device Philox draws (19 B), both through the shared-memory table
        try:
            n7 = 1 << 24
            g7 = torch.Generator(device=dev).manual_seed(11)
            in7 = [tuple(torch.randint(0, hi, (n7,), dtype=torch.uint8, device=dev, generator=g7) for hi in (5, 5, 16)) +
                   (torch.randint(-2**31, 2**31 - 1, (n7,), dtype=torch.int32, device=dev, generator=g7),) for _ in range(2)]
            variants = {}
            import numpy as np
            pol7 = np.random.RandomState(0).randint(0, 5, 761).astype(np.int8)
            # (tag, constructor kwargs, which inputs step() gets, algorithmic bytes per env-step)
            cases7 = (("slip0.2_injected", dict(slip_prob=0.2), "abr32", 24),
                      ("slip0.2_philox", dict(slip_prob=0.2, rng_mode="philox"), "ab", 19),
                      ("philox_draws", dict(rng_mode="philox"), "ab", 19),
                      ("single_agent_b_folded", dict(player_b_policy=pol7), "ar", 19),
                      ("single_agent_a_folded_philox", dict(player_a_policy=pol7, rng_mode="philox"), "b", 18))
            for tag, kw, use, nbytes in cases7:
                e7 = SoccerVecEnv(n7, device=dev, kernel="table", want_reset_obs=False, **kw)

                def step7(i):
                    a_, b_, r_, r32_ = in7[i % 2]
                    e7.step(a_ if "a" in use else None, b_ if "b" in use else None, r_ if "r" in use else None,
                            rng32=r32_ if "32" in use else None)
                e7.reset(in7[0][2] if e7.rng_mode == "injected" else None)
                for i in range(30):                 # play the population in
                    step7(i)
                q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                q0.record()
                for i in range(20):
                    step7(i)
                q1.record()
                torch.cuda.synchronize()
                ms7 = q0.elapsed_time(q1) / 20
                variants[tag] = {"env_steps_per_s": n7 / (ms7 * 1e-3), "bytes_per_env_step": nbytes,
                                 "hbm_gbs": n7 * nbytes / (ms7 * 1e-3) / 1e9, "frac_of_hbm_peak": n7 * nbytes / (ms7 * 1e-3) / 1e9 / peak}
                del e7
            # K2 with slip_prob = 0.2 (2^22 envs x K = 16): integer-threshold fast path, 4 envs per thread
            e8 = SoccerVecEnv(1 << 22, slip_prob=0.2, device=dev, kernel="table", rng_mode="philox")
            e8.reset()
            b8 = (torch.empty((16, 1 << 22), dtype=torch.int32, device=dev), torch.empty((16, 1 << 22), dtype=torch.float32, device=dev),
                  torch.empty((16, 1 << 22), dtype=torch.uint8, device=dev))
            e8.rollout(64, want_streams=False)
            for _ in range(2):
                e8.rollout(16, out=b8)
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            q0.record()
            for _ in range(6):
                e8.rollout(16, out=b8)
            q1.record()
            torch.cuda.synchronize()
            ms8 = q0.elapsed_time(q1) / 6
            extra["k2_slip0.2_2^22_envs_K16"] = {"env_steps_per_s": (1 << 26) / (ms8 * 1e-3),
                                                 "hbm_gbs_at_9.125B": (1 << 26) * 9.125 / (ms8 * 1e-3) / 1e9,
                                                 "frac_of_hbm_peak": (1 << 26) * 9.125 / (ms8 * 1e-3) / 1e9 / peak}
            del e8, b8
            extra["k1_variants_2^24_envs"] = variants
            del in7
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            extra["k1_variants_2^24_envs"] = {"error": repr(e)}
        n2, T2 = 4096, 1000
        e2 = SoccerVecEnv(n2, device=dev, kernel="auto", want_reset_obs=False)
        a, b, r = (torch.randint(0, hi, (T2, n2), dtype=torch.uint8, device=dev) for hi in (5, 5, 16))
        o2 = (torch.empty((T2, n2), dtype=torch.int32, device=dev), torch.empty((T2, n2), dtype=torch.float32, device=dev),
              torch.empty((T2, n2), dtype=torch.uint8, device=dev), None)
        e2.reset(r[0])
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def timed(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            s0.record()
            for _ in range(reps):
                fn()
            s1.record()
            torch.cuda.synchronize()
            return s0.elapsed_time(s1) * 1e3 / (reps * T2)

        rows = [(a[t], b[t], r[t], (o2[0][t], o2[1][t], o2[2][t], None)) for t in range(T2)]   # views made up front

        def per_step():
            for ra, rb, rr, ro in rows:
                e2.step(ra, rb, rr, out=ro)
        us_py = timed(per_step)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            per_step()
        us_graph = timed(graph.replay)
        fused = {}
        for kern in ("rules", "table"):
            ef = SoccerVecEnv(n2, device=dev, kernel=kern, want_reset_obs=False)
            ef.reset(r[0])
            fused[kern] = timed(lambda: ef.step_many(a, b, r, out=o2), reps=10)
        us_fused = min(fused.values())
        # the same 4096 envs stepped from HOST buffers, closed loop (every step waits for its results)
        eh = SoccerVecEnv(n2, device=dev, kernel="auto", want_reset_obs=False)
        eh.reset(r[0])
        h3 = [tuple(x[t].cpu().pin_memory() for x in (a, b, r)) for t in range(8)]
        host_us = {}
        for name, zc in (("zero_copy", True), ("staged", False)):
            for t in range(20):
                eh.step_host(*h3[t % 8], zero_copy=zc)
            t0 = time.perf_counter()
            for t in range(500):
                eh.step_host(*h3[t % 8], zero_copy=zc)
            host_us[name] = (time.perf_counter() - t0) / 500 * 1e6
        ep = SoccerVecEnv(n2, device=dev, kernel="table", want_reset_obs=False)
        ep.reset(r[0])
        hp = []
        for t in range(8):
            jb, rb = ep.alloc_host_inputs(packed=True)
            jb.copy_(SoccerVecEnv.pack_joint(h3[t][0], h3[t][1])); rb.copy_(h3[t][2])
            hp.append((jb, rb))
        for t in range(20):
            ep.step_host_packed(*hp[t % 8])
        t0 = time.perf_counter()
        for t in range(500):
            ep.step_host_packed(*hp[t % 8])
        host_us["packed_zero_copy_table_kernel"] = (time.perf_counter() - t0) / 500 * 1e6
        del ep, hp
        extra["config2_4096_envs"] = {
            "us_per_step_host_buffers_closed_loop": host_us,
            "k1_kernel": e2.kernel, "us_per_step_k1_python_loop": us_py, "us_per_step_k1_cuda_graph": us_graph,
            "us_per_step_fused_replay": fused, "env_steps_per_s": n2 / (us_fused * 1e-6),
            "note": "82 KB per step: K1 is launch-latency bound here (not graded against the HBM roofline); "
                    "soccer_step_many runs the T = 1000 steps in ONE launch with the state in registers"}
        # the same fused replay at 2^22 envs x 64 steps: 12.125 B / env-step instead of K1's 20
        n5, T5 = 1 << 22, 64
        a, b, r = (torch.randint(0, hi, (T5, n5), dtype=torch.uint8, device=dev) for hi in (5, 5, 16))
        o5 = (torch.empty((T5, n5), dtype=torch.int32, device=dev), torch.empty((T5, n5), dtype=torch.float32, device=dev),
              torch.empty((T5, n5), dtype=torch.uint8, device=dev), None)
        T2 = T5
        big = {}
        for kern in ("rules", "table"):
            ef = SoccerVecEnv(n5, device=dev, kernel=kern, want_reset_obs=False)
            ef.reset(r[0])
            us = timed(lambda: ef.step_many(a, b, r, out=o5), reps=8)
            big[kern] = {"env_steps_per_s": n5 / (us * 1e-6), "hbm_gbs_at_12.125B": n5 * 12.125 / (us * 1e-6) / 1e9,
                         "frac_of_hbm_peak": n5 * 12.125 / (us * 1e-6) / 1e9 / peak}
        extra["fused_replay_2^22_envs_T64"] = big
        # the reference's value iteration (planners.py:4-18) as ONE cooperative kernel launch
        from gym_soccer_littman94_b200.utils import planners
        from gym_soccer_littman94_b200.utils.policies import get_random_policy
        ev = SoccerSimultaneousEnv(5, 4, slip_prob=0.2, player_b_policy=get_random_policy(761, 5, seed=42), device=dev)
        planners.value_iteration(ev, 1e-10, 0.99)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, _, _, vi_cc = planners.value_iteration(ev, 1e-10, 0.99)
        vi_ms = (time.perf_counter() - t0) * 1e3
        extra["value_iteration_5x4_slip0.2"] = {
            "sweeps": vi_cc, "ms": vi_ms, "us_per_sweep": vi_ms * 1e3 / vi_cc,
            "note": "soccer_plan: whole value iteration (theta 1e-10, gamma 0.99) in one cooperative launch, bit-identical "
                    "to the reference's planner, which takes 5.6 s for 183 sweeps on this env class in the build "
                    "container (tests/golden/ref_planner_5x4_s020_a_free.npz vi_seconds)"}
    except Exception as e:  # noqa: BLE001
        extra["error"] = repr(e)

    cores = os.cpu_count() or 1
    cpu_v, cpu_sample = time_cpu_port(args.cpu_seconds, 1)
    cpu_all_v, _ = time_cpu_port(min(args.cpu_seconds, 5.0), cores)
    # BASELINE config 1 through the UNMODIFIED Python reference on this host (oracle/_ref, staged by build()): 100,000
    # step() calls, 1 core, RandomState(123) actions -- BASELINE.md's CPU-baseline plan
    py_ref = None
    try:
        from oracle import time_reference
        py_ref = time_reference.time_python_reference(100000, 0.0)
        if py_ref is not None and "error" not in py_ref:
            py_ref["slip_0_2_steps_per_s"] = (time_reference.time_python_reference(50000, 0.2) or {}).get("steps_per_s")
    except Exception as e:  # noqa: BLE001
        py_ref = {"error": repr(e)}

    kname = "k_step_table" if env.kernel == "table" else "k_step_fast"
    # the timed launches are the STATS instantiations (statistics fused into the step)
    kexact = "k_step_table<0, 0, 0, 0, 1>" if env.kernel == "table" else "k_step_fast<0, 0, 0>"
    traffic, traffic_src = ncu_traffic(kname, kexact) if N == (1 << 24) else (None, None)
    narrow_api = ("SoccerVecEnv.step_host(narrow=True): uint8 act_a / act_b / draws up; obs uint16, reward int8, flags "
                  "uint8 down; every step waits for its results")
    e2e_narrow = {"value": e2e_value, "h2d_bytes_per_step": 3 * N, "d2h_bytes_per_step": 4 * N, "checksum": checksum,
                  "chunks": args.chunks, "api": narrow_api}
    e2e_wide = {"value": e2e[False][0], "h2d_bytes_per_step": 3 * N, "d2h_bytes_per_step": 9 * N,
                "api": "step_host(narrow=False): obs int32, reward float32, flags uint8 down"}
    # headline e2e: the fastest host-buffer call of the public API that carries the full step result (packed streams
    # on the table kernel); the natural-dtype formats are reported next to it
    if e2e_packed is not None:
        e2e_line = dict(e2e_packed, unit=UNIT, steps=Ke, cpus_bound_near_gpu=numa_cpus,
                        closed_loop="every step waits for its results on the host before the next one is enqueued",
                        narrow=e2e_narrow, wide=e2e_wide, slip_0_2=e2e_slip)
    else:
        e2e_line = dict(e2e_narrow, unit=UNIT, steps=Ke, cpus_bound_near_gpu=numa_cpus, wide=e2e_wide)
    emit_json(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/u32 integer (f32 reward stream)", "data": "synthetic",
        "config": {"workload": "K1 streaming lock-step step: 2^24 5x4 envs per GPU, host-supplied joint actions, "
                               "injected 2-bit draws, fused auto-reset (BASELINE configs[1] semantics at the N "
                               "BASELINE.md grades the roofline on)",
                   "envs_per_gpu": N, "kernel": env.kernel, "pitch": "5x4", "slip_prob": 0.0,
                   "l2": f"per-step traffic {BYTES_PER_ENV_STEP * N / 1e6:.0f} MB > 126 MB L2; inputs and outputs "
                         f"cycle through {RING}-deep rings", "parallelism": f"independent env shards x{world}",
                   "population": "played in for 128 untimed steps after reset() (a fresh reset puts all envs on 4 observations)",
                   "statistics": "episode statistics accumulated inside the step kernel at every N"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "traffic_note": "DRAM bytes per launch in an "
                     "isolated (ncu-serialised, cache-flushed) launch; below the algorithmic 335.5 MB because dirty lines "
                     "still sit in L2 when the kernel ends",
                     "l2_note": "achieved counts ALGORITHMIC bytes: at 2^24 envs the 67 MB state tensor written by step i "
                                "is partly still in the 126 MB L2 when step i+1 reads it, so frac can exceed the DRAM-only "
                                "figure (extra.k1_2^26_envs: 268 MB of state, no reuse)",
                     "peak_source": peak_src, "kernel": kname,
                     "bytes_per_env_step": BYTES_PER_ENV_STEP, "avg_launch_ms": kern_ms, "min_launch_ms_event_after_every_launch": min(per_launch),
                     "stream_mix_probe": {"achieved": probe_gbs, "unit": "GB/s", "kernel_over_probe": achieved / probe_gbs,
                                          "what": "k_stream_mix_probe: K1's streams, access pattern and launch shape "
                                                  "without the game logic = practical ceiling for its 7 B read / "
                                                  "13 B written mix"}},
        "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": 1, "kind": "port", "sample": cpu_sample, "host": host_info(),
                         "all_cores": {"value": cpu_all_v, "cores": cores},
                         "python_reference": py_ref if py_ref is not None else
                         "oracle/_ref absent (build() stages it where /root/reference exists)"},
        "host": host_info(),
        "shard_invariant": shard.get("shard_invariant"), "shard_invariant_detail": shard,
        "timed_region": {"steps": K, "statistics": "fused into K1 (soccer_step_args.stats)", "tail_after_last_step_us": tail_us, "tail_us_per_rank": tails,
                         "tail_is": "the sum all-reduce of the 48-byte statistics vector + waiting for the slowest rank"
                         if world > 1 else "nothing (1 GPU)", "collective": collective},
        "e2e": e2e_line,
        "sustained": sustained,
        "gpu_launches": K + (1 if p2p is not None else 0),      # K step kernels (+ the peer-memory all-reduce kernel at N > 1)
        "clocks": clocks,
        "extra": extra,
        "stats_allreduce": stats_timed,
    }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 24)
    ap.add_argument("--kernel", default="auto", choices=["auto", "rules", "table"])
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--chunks", type=int, default=8, help="pipeline slices of the host-buffer step")
    ap.add_argument("--no-bind", action="store_true", help="do not pin ranks to the CPU cores local to their GPU")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"],
                    help="statistics all-reduce: the peer-memory kernel (default; NCCL if unavailable) or NCCL")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything native libraries print there while the bench runs (NCCL's
    # "NCCL version ..." banner, for one) is sent to stderr instead
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    _JSON_OUT.flush()


if __name__ == "__main__":
    main()
