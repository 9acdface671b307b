"""Multi-GPU plumbing: one process per GPU, environments sharded by contiguous global id range.

Environments never exchange state, so the data path has NO collective: each rank
steps its own shard with `env_id_base = first global id of the shard`, and because the Philox
draws are keyed by the GLOBAL env id the trajectories are identical at 1, 2, 4 or 8 GPUs.  The
one collective of the path is a sum all-reduce of the 6-entry int64 episode-statistics vector
[episodes, goals_A, goals_B, truncations, steps, sum_episode_len]: on GPUs one hand-written kernel
over NVLink peer memory (P2PStatsAllReduce, soccer_stats_allreduce_p2p), with the process group's own
all-reduce as the fallback (NCCL; gloo with CPU tensors in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist

STAT_NAMES = ("episodes", "goals_A", "goals_B", "truncations", "steps", "sum_episode_len")


def shard_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(env_id_base, n_local) of `rank`: contiguous, sizes differ by at most one, and every shard
    but possibly the last few is a multiple of 4 envs so the 128-bit paths stay aligned."""
    assert 0 <= rank < world_size and total_envs >= 0
    groups = (total_envs + 3) // 4                       # 4-env groups, the kernels' vector unit
    per, extra = divmod(groups, world_size)
    g0 = rank * per + min(rank, extra)
    g1 = g0 + per + (1 if rank < extra else 0)
    lo, hi = min(4 * g0, total_envs), min(4 * g1, total_envs)
    return lo, hi - lo


class P2PStatsAllReduce:
    """The statistics all-reduce as ONE kernel over NVLink peer memory (soccer_stats_allreduce_p2p): a symmetric
    256-slot buffer per rank (torch.distributed._symmetric_memory provides the cross-process mapping -- plumbing),
    peer stores + a release / acquire epoch flag, enqueued on the caller's stream right behind the last step kernel.
    `P2PStatsAllReduce.create()` returns None where symmetric memory cannot be set up (no peer access, CPU tensors,
    a single rank): the callers then fall back to the NCCL / gloo all-reduce."""

    def __init__(self, device):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self._C, self._lib = C, _lib.lib()
        nbytes = C.c_int64()
        _lib.check(self._lib.soccer_stats_allreduce_p2p_bytes_host(C.byref(nbytes)), "p2p bytes")
        self.buf = symm_mem.empty(nbytes.value // 8, dtype=torch.int64, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, dist.group.WORLD)
        self.rank, self.world = int(self.hdl.rank), int(self.hdl.world_size)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        assert len(ptrs) == self.world <= 16
        self.ptrs = (C.c_uint64 * self.world)(*ptrs)
        self.epoch = 0
        self.device = device
        torch.cuda.synchronize(device)
        dist.barrier()                       # every buffer is zero-filled before anybody stores into it

    @classmethod
    def create(cls, device):
        try:
            if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
                return None
            if torch.device(device).type != "cuda" or dist.get_backend() != "nccl":
                return None
            ok = torch.zeros(1, dtype=torch.int32, device=device)
            try:
                obj = cls(device)
                ok += 1
            except Exception:  # noqa: BLE001
                obj = None
            dist.all_reduce(ok)              # all ranks or none
            return obj if int(ok.item()) == dist.get_world_size() else None
        except Exception:  # noqa: BLE001
            return None

    def __call__(self, stats: torch.Tensor) -> torch.Tensor:
        C = self._C
        assert stats.is_cuda and stats.dtype == torch.int64 and stats.is_contiguous() and stats.numel() >= 6
        self.epoch += 1
        from ._lib import check
        check(self._lib.soccer_stats_allreduce_p2p(self.ptrs, self.rank, self.world, self.epoch, C.c_void_p(stats.data_ptr()),
                                                   C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
              "soccer_stats_allreduce_p2p")
        return stats


def allreduce_stats(stats: torch.Tensor, p2p: "P2PStatsAllReduce | None" = None) -> torch.Tensor:
    """Sum the per-rank statistics vector over all ranks (in place); a no-op without a group.  p2p: the peer-memory
    kernel (P2PStatsAllReduce.create(device)); None = the process group's all-reduce (NCCL on GPU tensors, gloo on CPU)."""
    assert stats.dtype == torch.int64 and stats.numel() == len(STAT_NAMES)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if p2p is not None:
            return p2p(stats)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def stats_dict(stats: torch.Tensor) -> dict:
    return {k: int(v) for k, v in zip(STAT_NAMES, stats.cpu().tolist())}


class ShardedSoccerVecEnv:
    """The global batch of `total_envs` environments, of which this rank owns one shard."""

    def __init__(self, total_envs: int, rank: int = None, world_size: int = None, device=None, **env_kwargs):
        from .envs import SoccerVecEnv
        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
        if world_size is None:
            world_size = dist.get_world_size() if dist.is_initialized() else 1
        self.rank, self.world_size, self.total_envs = rank, world_size, total_envs
        self.env_id_base, self.n_local = shard_range(total_envs, rank, world_size)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.env = SoccerVecEnv(self.n_local, device=device, env_id_base=self.env_id_base, **env_kwargs)
        self.p2p = P2PStatsAllReduce.create(device) if torch.device(device).type == "cuda" else None

    def reset(self, *a, **kw):
        return self.env.reset(*a, **kw)

    def step(self, *a, **kw):
        return self.env.step(*a, **kw)

    def rollout(self, K: int, **kw):
        """Local fused rollout + the global statistics (one all-reduce of 48 bytes)."""
        obs, reward, flags, stats = self.env.rollout(K, **kw)
        return obs, reward, flags, allreduce_stats(stats, self.p2p)
