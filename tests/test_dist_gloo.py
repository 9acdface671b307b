"""world_size-2 (and 3) CPU tests of the multi-GPU host logic over gloo: shard partition,
global-env-id keying and the one statistics all-reduce.  The per-shard transitions here come from
the CPU oracle (test infrastructure) -- what is under test is the sharding / reduction plumbing
in gym_soccer_littman94_b200/dist.py, which is the code the NCCL path runs too."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gym_soccer_littman94_b200.dist import STAT_NAMES, allreduce_stats, shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 3, 4, 5, 1023, 4096, (1 << 24) + 7):
        for world in (1, 2, 3, 4, 8):
            pos = 0
            for r in range(world):
                lo, n = shard_range(total, r, world)
                assert lo == pos and n >= 0
                if r < world - 1 and n:
                    assert lo % 4 == 0                      # every shard starts 128-bit aligned
                pos += n
            assert pos == total
    assert shard_range(1 << 24, 3, 8) == (3 << 21, 1 << 21)      # BASELINE config 4: 2^21 per GPU


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, K, seed, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import soccer_oracle as so
    m = so.OracleModel(5, 4, 0.0)
    base, n = shard_range(total, rank, world)
    states = np.zeros(n, so.STATE_DTYPE)
    for i in range(n):       # initial reset keyed by the GLOBAL env id, like SoccerVecEnv.reset (philox mode)
        states[i] = m.isd[so.philox_decode(so.philox_word(seed, base + i, (1 << 64) - 1))[3]][1]
    ts = np.zeros(n, np.int32)
    obs, rew, flg, st = m.rollout_philox(states, ts, K, seed, env_id_base=base)
    stats = torch.from_numpy(st.copy())
    allreduce_stats(stats)                                  # the one collective of the path
    gathered = [None] * world
    dist.all_gather_object(gathered, (base, obs))
    if rank == 0:
        out_q.put((stats.numpy().copy(), gathered))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_rollout_equals_single_process(world):
    """Results are independent of the number of ranks: concatenated shard streams == the
    single-process rollout, and the all-reduced statistics == the single-process statistics."""
    from oracle import soccer_oracle as so
    total, K, seed = 1001, 40, 77
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, K, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    stats, gathered = q.get()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    m = so.OracleModel(5, 4, 0.0)
    states = np.zeros(total, so.STATE_DTYPE)
    for i in range(total):
        states[i] = m.isd[so.philox_decode(so.philox_word(seed, i, (1 << 64) - 1))[3]][1]
    obs, rew, flg, st = m.rollout_philox(states, np.zeros(total, np.int32), K, seed)
    assert np.array_equal(stats, st)
    assert dict(zip(STAT_NAMES, st))["steps"] == total * K
    full = np.concatenate([o for _, o in sorted(gathered, key=lambda t: t[0])], axis=1)
    assert np.array_equal(full, obs)
