"""world_size-2 gloo run of the N>1 host logic: shard ranges, env offsets and the episode-metric all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import swarmacb_isaaclab_b200 as pkg
from swarmacb_isaaclab_b200.sharding import METRIC_NAMES, EpisodeMetrics, shard_cfg, shard_range


def test_shard_ranges_partition_the_job():
    for total in (0, 1, 7, 16384, 65536, 65537):
        for world in (1, 2, 3, 8):
            ranges = [shard_range(total, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_shard_cfg_sets_envs_and_offset():
    cfg = pkg.ForagingEnvCfg()
    cfg.update_variant("daisy")
    local, off = shard_cfg(cfg, 65536, 8, 3, device="cuda:3")
    assert (local.scene.num_envs, off, local.sim.device) == (8192, 3 * 8192, "cuda:3")
    assert cfg.scene.num_envs == 5  # the caller's cfg is untouched


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, stop = shard_range(10, world, rank)
    E = stop - start
    m = EpisodeMetrics("cpu")
    reward = torch.full((E,), float(rank + 1))
    time_out = torch.zeros(E, dtype=torch.bool)
    time_out[0] = True
    m.update(reward, time_out, torch.full((E,), 10.0 * (rank + 1)), max_episode_length=1200, n_agents=20)
    out = m.reduce()
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, out))


def test_metric_all_reduce_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = {"sum_episode_return": 10.0 + 20.0, "sum_episode_length": 2400.0, "n_episodes": 2.0,
            "sum_group_reward": 5 * 1.0 + 5 * 2.0, "agent_steps": 10 * 20.0}
    for r in (0, 1):
        assert set(results[r]) == set(METRIC_NAMES)
        for k, v in want.items():
            assert results[r][k] == v, (r, k, results[r][k], v)
