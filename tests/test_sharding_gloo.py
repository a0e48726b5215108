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


def _clock_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import fixtures
    from oracle_env import OracleBackedEnv
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = fixtures.make_cfg("xor", "cyclamen", 4)
    cfg.episode_length_s = 1.0                       # L = 10 steps
    env = OracleBackedEnv(cfg, env_offset=4 * rank)
    env.reset()
    # rank 0: counters 0,0,3,3 -> roll-overs 9 and 6 steps from now; rank 1: 0,7,7,9 -> 9, 2 and 0 steps from now
    env.episode_length_buf = torch.tensor([0, 0, 3, 3] if rank == 0 else [0, 7, 7, 9])
    clock = env.attach_job_reset_clock()
    s0 = env._step_counter
    bits = clock.bits(s0, 12)
    seen = []
    for t in range(12):   # the flag each step would be handed, and whether one of MY envs really timed out
        _, _, to = env.step_tensor(torch.zeros(4, 20, 1, dtype=torch.long))
        seen.append(bool(to.any()))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bits, seen))


def test_job_reset_clock_is_job_wide_gloo_world2():
    """The ENV:1262 any-reset flag of a sharded job: both ranks derive the SAME schedule, the OR of both shards'
    roll-over phases, from one all-reduce; it matches the time-outs that then really happen."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_clock_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r: (b, s) for r, b, s in (q.get(timeout=180) for _ in procs)}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_steps = {0, 2, 6, 9, 10}                      # phases {9, 6} | {9, 2, 0}, period 10, within 12 steps
    want_bits = sum(1 << t for t in want_steps | {t + 10 for t in want_steps if t + 10 < 12})
    assert res[0][0] == res[1][0] == want_bits
    union = [a or b for a, b in zip(res[0][1], res[1][1])]
    assert [t for t, hit in enumerate(union) if hit] == sorted(t for t in range(12) if (want_bits >> t) & 1)
