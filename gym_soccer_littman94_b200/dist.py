"""Multi-GPU plumbing: one process per GPU, environments sharded by contiguous global id range.

Environments never exchange state, so the data path has NO collective: each rank
steps its own shard with `env_id_base = first global id of the shard`, and because the Philox
draws are keyed by the GLOBAL env id the trajectories are identical at 1, 2, 4 or 8 GPUs.  The
one collective of the path is a sum all-reduce of the 6-entry int64 episode-statistics vector
[episodes, goals_A, goals_B, truncations, steps, sum_episode_len] (NCCL on GPU tensors; the
same code runs over gloo with CPU tensors in the CPU tests).
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


def allreduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """Sum the per-rank statistics vector over all ranks (in place); a no-op without a group."""
    assert stats.dtype == torch.int64 and stats.numel() == len(STAT_NAMES)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
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

    def reset(self, *a, **kw):
        return self.env.reset(*a, **kw)

    def step(self, *a, **kw):
        return self.env.step(*a, **kw)

    def rollout(self, K: int, **kw):
        """Local fused rollout + the global statistics (one all-reduce of 48 bytes)."""
        obs, reward, flags, stats = self.env.rollout(K, **kw)
        return obs, reward, flags, allreduce_stats(stats)
