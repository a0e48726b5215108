"""BASELINE config 1 on the GPU: the fused manual-control tick against the reference's StandaloneDGTEnv rollouts
(teacher-forced tick by tick, full 1800-tick XOR trajectory) and against the oracle bit for bit."""
import numpy as np
import pytest
import torch

import fixtures
from oracle import oracle
from swarmacb_isaaclab_b200.params import N

pytestmark = pytest.mark.gpu
FILES = fixtures.mc_fixture_files()


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1][:-4])
def test_cuda_mc_tick_matches_manual_control_rollout(path):
    from swarmacb_isaaclab_b200.standalone import StandaloneSwarmEnv
    fx = fixtures.McFixture(path)
    env = StandaloneSwarmEnv(task=fx.meta["task"], device="cuda:0")
    stride = 1 if fx.ticks <= 200 else 3      # the 1800-tick rollout is checked on every 3rd tick (+ the roll-over)
    ticks = sorted(set(range(0, fx.ticks, stride)) | set(np.nonzero(fx.z["rolled"])[0].tolist()))
    for t in ticks:
        env.load_state(fx.state(t))
        inp = fx.tick_inputs(t)
        env.inject_noise(rab_u=inp["rab_u"], rab_u2=inp["rab_u2"], turn_dur=inp["turn_dur"], mc_spawn_u=inp["mc_spawn_u"])
        obs, reward, rolled = env.tick(inp["module_ids"], float(inp["wheels"][0, 0, 0]), float(inp["wheels"][0, 0, 1]))
        torch.cuda.synchronize()
        fx.check(t, env.dump_state() | {"completed_group_reward": None}, obs.cpu().numpy(), reward.cpu().numpy(),
                 rolled.cpu().numpy(), label=f"cuda:{fx.name}")


@pytest.mark.parametrize("task", ["SwarmACB-XOR-v0", "SwarmACB-Foraging-v0", "SwarmACB-Sheltering-v0", "SwarmACB-DirectionalGate-v0"])
def test_cuda_mc_tick_equals_oracle(task):
    """Free-running batch of standalone envs vs the oracle, bit for bit (poses, FSM, rewards, observations)."""
    from swarmacb_isaaclab_b200.standalone import StandaloneSwarmEnv
    E, T = 32, 40
    rng = np.random.default_rng(3)
    env = StandaloneSwarmEnv(task=task, device="cuda:0", num_envs=E)
    p = env.params
    host = oracle.new_state(E)
    spawn = rng.random((E, N, 3), dtype=np.float32)
    spawn[: E // 2, :, 0] *= 0.02                     # half of the envs start as one tight cluster
    env.inject_noise(mc_spawn_u=spawn)
    env.reset()
    oracle.mc_reset(p, host, spawn)
    host["episode_length_buf"][::3] = p.max_episode_length - 7     # roll-overs inside the window
    env.step_count.copy_(torch.as_tensor(host["episode_length_buf"]))
    for t in range(T):
        ids = rng.integers(0, 6, (E, N), dtype=np.int64)
        wheels = np.zeros((E, N, 2), np.float32)
        wheels[:, 0] = rng.random((E, 2), dtype=np.float32) * 0.4 - 0.2
        noise = dict(rab_u=rng.random((E, N, N), dtype=np.float32), rab_u2=rng.random((E, N, N), dtype=np.float32),
                     turn_dur=rng.integers(1, 5, (E, N, 3)).astype(np.int32), mc_spawn_u=rng.random((E, N, 3), dtype=np.float32))
        env.inject_noise(**noise)
        obs, rew, rolled = env.tick(ids, wheels[:, 0, 0], wheels[:, 0, 1])
        obs_o, rew_o, rolled_o = oracle.mc_tick(p, host, ids, wheels, **noise)
        torch.cuda.synchronize()
        dev = env.dump_state()
        for k in dev:
            assert np.array_equal(dev[k], host[k]), f"{task} t={t}: {k}"
        assert np.array_equal(obs.cpu().numpy(), obs_o), f"{task} t={t}: obs"
        assert np.array_equal(rew.cpu().numpy(), rew_o) and np.array_equal(rolled.cpu().numpy(), rolled_o)
    assert rolled_o.dtype == bool
