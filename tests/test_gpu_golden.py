"""GPU parity: the fused CUDA step (through the C ABI / SwarmEnv) against the golden fixtures."""
import numpy as np
import pytest
import torch

import fixtures

pytestmark = pytest.mark.gpu
FILES = fixtures.fixture_files()


def _env(fx):
    from swarmacb_isaaclab_b200.env import SwarmEnv
    cfg = fixtures.make_cfg(fx.meta["mission"], fx.meta["mode"], fx.E, fx.meta["decimation"], device="cuda:0")
    return SwarmEnv(cfg)


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1][:-4])
def test_cuda_step_matches_reference(path):
    fx = fixtures.Fixture(path)
    env = _env(fx)
    for t in range(fx.steps):
        case = fx.step_case(t)
        env.load_state(case["pre"])
        env.inject_noise(rab_u=case["rab_u"], turn_dur=case["turn_dur"], spawn_u=case["spawn_u"], yaw_u=case["yaw_u"])
        act = torch.as_tensor(case["actions"], device=env.device)
        obs, reward, time_out = env.step_tensor(act)
        critic = env.get_critic_state()
        torch.cuda.synchronize()
        fixtures.compare(case, fx.params, env.dump_state(), obs.cpu().numpy(), reward.cpu().numpy(),
                         time_out.cpu().numpy(), critic.cpu().numpy(), label=f"cuda:{fx.name}[{t}]")
