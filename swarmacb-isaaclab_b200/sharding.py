"""Environment sharding across GPUs.

Environments are independent, so a job of ``total_envs`` is cut into contiguous ranges, one per
rank (one process per GPU).  The step has NO collective; the only exchange is a latency-bound
all-reduce(sum) of a five-float episode-metric vector at report time (NCCL over NVLink on GPUs, gloo
in the CPU tests).  The in-kernel Philox stream is keyed by the GLOBAL env index
(``env_offset + local index``), so a trajectory does not depend on how many GPUs the job uses.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

METRIC_NAMES = ("sum_episode_return", "sum_episode_length", "n_episodes", "sum_group_reward", "agent_steps")


def shard_range(total_envs: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous [start, stop) env range of ``rank``; the first ``total % world`` ranks get one extra env."""
    if total_envs < 0 or world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError(f"bad shard request total={total_envs} world={world_size} rank={rank}")
    base, extra = divmod(total_envs, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_cfg(cfg, total_envs: int, world_size: int, rank: int, device: str | None = None):
    """Return (copy of cfg for this rank with scene.num_envs set, env_offset for SwarmEnv)."""
    start, stop = shard_range(total_envs, world_size, rank)
    local = cfg.copy()
    local.scene.num_envs = stop - start
    if device is not None:
        local.sim.device = device
    return local, start


class EpisodeMetrics:
    """Accumulates per-shard episode statistics on the device and reduces them over all ranks."""

    def __init__(self, device):
        self.vec = torch.zeros(len(METRIC_NAMES), dtype=torch.float64, device=device)

    def update(self, reward: torch.Tensor, time_out: torch.Tensor, completed_group_reward: torch.Tensor,
               max_episode_length: int, n_agents: int):
        """Fold one step's outputs in: rewards (E,), time_out (E,) bool, completed_group_reward (E,)."""
        done = time_out.to(torch.float64)
        n_done = done.sum()
        self.vec[0] += (completed_group_reward.to(torch.float64) * done).sum()
        self.vec[1] += n_done * max_episode_length
        self.vec[2] += n_done
        self.vec[3] += reward.to(torch.float64).sum()
        self.vec[4] += reward.numel() * n_agents

    def reduce(self, group=None) -> dict:
        """All-reduce(sum) over the job; returns a name -> float dict (host sync)."""
        out = self.vec.clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return dict(zip(METRIC_NAMES, out.tolist()))
