"""Environment sharding across GPUs.

Environments are independent, so a job of ``total_envs`` is cut into contiguous ranges, one per
rank (one process per GPU).  The step has NO collective; the only exchange is a latency-bound
all-reduce(sum) of a five-float episode-metric vector at report time (NCCL over NVLink on GPUs, gloo
in the CPU tests).  The in-kernel Philox stream is keyed by the GLOBAL env index
(``env_offset + local index``), so a trajectory does not depend on how many GPUs the job uses -

with one caveat the reference itself carries: ``_reset_idx`` re-solves the collisions of ALL environments of the
batch whenever ANY of them times out (ENV:1262).  A shard sees only its own environments, so that coupling is per
shard by default - exactly what the reference does when a job runs as independent processes of E/G environments.
While all episode counters move in lockstep (the normal case: every env starts at 0) every shard rolls over at the
same step and sharded == unsharded bit for bit.  With de-synchronised counters (a resumed checkpoint, counters
written from outside) attach a :class:`JobResetClock`: one all-reduce of an L-bit phase mask when it is built, no
collective per step, and every shard is told the job-wide flag (``SwarmNoise.any_reset_mode``).
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


class JobResetClock:
    """Which upcoming steps see a time-out SOMEWHERE in the job (ENV:1262 couples the whole batch).

    Episodes only ever end by time-out (``terminated`` is always False, ENV:1200-1209), so env e rolls over at the
    steps ``k = (L - 1 - len_e) mod L`` (mod L): the job-wide set of roll-over phases is an L-bit mask.  It is built
    ONCE - per-rank mask from the rank's counters, OR-ed over the ranks with one all-reduce - and then answers every
    step from the host without touching the device.  It must be rebuilt (collectively) after episode_length_buf is
    written from outside or a checkpoint is loaded; ``SwarmEnv`` does that itself when a clock is attached.
    """

    def __init__(self, env, group=None, peers=None):
        """``group``: the process group of the job (one shard per rank).  ``peers``: further shards living in THIS
        process (several envs on one GPU, or one per stream) whose counters belong to the same job."""
        self.group = group
        self.peers = list(peers) if peers else []
        self.rebuild(env)

    def rebuild(self, env):
        L = int(env.max_episode_length)
        phase = torch.zeros(L, dtype=torch.int32, device=env.episode_length_buf.device)
        for shard in [env] + [p for p in self.peers if p is not env]:
            lens = shard.episode_length_buf.to(torch.long)
            # env e times out at the step where its counter reads L - 1, i.e. (L - 1 - len_e) steps from now
            phase[(L - 1 - lens.clamp(0, L - 1)) % L] = 1
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(phase, op=dist.ReduceOp.MAX, group=self.group)
        self.phase = phase.cpu().numpy().astype(bool)
        self.period = L
        self.origin = int(env._step_counter)

    def bits(self, step_counter: int, n_steps: int = 1) -> int:
        """Bit t set = some env of the job times out at step ``step_counter + t``."""
        out = 0
        for t in range(n_steps):
            if self.phase[(step_counter + t - self.origin) % self.period]:
                out |= 1 << t
        return out
