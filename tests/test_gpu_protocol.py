"""GPU checks of the task API the reference's trainers drive: the dict protocol at the trainers'
cadence, get_critic_state() fused into the step (ABI v2), completed_terminal_critic_state, attribute rebinding."""
import numpy as np
import pytest
import torch

import fixtures
from oracle import oracle
from swarmacb_isaaclab_b200 import _lib
from swarmacb_isaaclab_b200.params import N

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _mk(mission, mode, E, **kw):
    from swarmacb_isaaclab_b200.env import SwarmEnv
    return SwarmEnv(fixtures.make_cfg(mission, mode, E, 1, device=DEV), **kw)


def _actions(rng, p, E):
    if p.discrete_actions:
        return rng.integers(0, 6, (E, N, 1), dtype=np.int64)
    return (rng.random((E, N, 2), dtype=np.float32) * 2 - 1).astype(np.float32)


@pytest.mark.parametrize("mission,mode", [("shl", "oc2"), ("for", "daisy"), ("dgt", "dandelion")])
def test_trainer_style_decisions_match_oracle(mission, mode):
    """What collect_rollout does per decision (agents/poca_trainer.py:509-583): get_critic_state(), an action dict of
    20 strided views, decision_period env.steps with that dict, then completed_terminal_critic_state - on the GPU
    through the dict API, every step checked against the oracle (some envs time out inside the window)."""
    E, decisions, period = 96, 3, 5
    rng = np.random.default_rng(11)
    env = _mk(mission, mode, E)
    p = env.params
    host = oracle.new_state(E)
    spawn_u, yaw_u = rng.random((6, E, N, 2), dtype=np.float32), rng.random((E, N), dtype=np.float32)
    rab_u = rng.random((E, N, N), dtype=np.float32)
    env.inject_noise(rab_u=rab_u, spawn_u=spawn_u, yaw_u=yaw_u)
    obs_dict, _ = env.reset()
    obs_o = oracle.reset(p, host, rab_u=rab_u, spawn_u=spawn_u, yaw_u=yaw_u)
    assert np.array_equal(torch.stack([obs_dict[a] for a in env.possible_agents], 1).cpu().numpy(), obs_o)
    # some envs time out at the 2nd / 7th / 12th step of the window
    for k, off in ((0, 2), (1, 7), (2, 12)):
        host["episode_length_buf"][k::9] = p.max_episode_length - off
    env.episode_length_buf = torch.as_tensor(host["episode_length_buf"])   # rebinding must reach the kernel's tensor
    lib = _lib.load()
    for d in range(decisions):
        launches = lib.swarm_kernel_launch_count()
        crit = env.get_critic_state()
        if d >= 2:  # cadence learned after two calls: the previous step wrote it, nothing is launched here
            assert lib.swarm_kernel_launch_count() == launches
        assert np.abs(crit.cpu().numpy() - oracle.critic_state(p, host)).max() <= 2e-5, f"critic d={d}"
        crit_before = crit.clone()
        act = _actions(rng, p, E)
        act_t = torch.as_tensor(act, device=DEV)
        action_dict = {a: act_t[:, i] for i, a in enumerate(env.possible_agents)}
        for t in range(period):
            rab_u = rng.random((E, N, N), dtype=np.float32)
            dur = rng.integers(1, 5, (E, N, 3)).astype(np.int32)
            spawn_u, yaw_u = rng.random((6, E, N, 2), dtype=np.float32), rng.random((E, N), dtype=np.float32)
            env.inject_noise(rab_u=rab_u, turn_dur=dur, spawn_u=spawn_u, yaw_u=yaw_u)
            obs_dict, rew, term, trunc, _ = env.step(action_dict)
            obs_o, rew_o, to_o = oracle.step(p, host, act.reshape(E, N, -1).squeeze(-1) if p.discrete_actions else act,
                                             rab_u=rab_u, turn_dur=dur, spawn_u=spawn_u, yaw_u=yaw_u)
            lab = f"{mission}/{mode} d={d} t={t}"
            a0 = env.possible_agents[0]
            assert np.array_equal(rew[a0].cpu().numpy(), rew_o), lab
            assert np.array_equal(trunc[a0].cpu().numpy(), to_o), lab
            assert not bool(term[a0].any()), lab
            got = torch.stack([obs_dict[a] for a in env.possible_agents], 1).cpu().numpy()
            assert np.array_equal(got, obs_o), lab
            dev = env.dump_state()
            assert np.array_equal(dev["pos"], host["pos"]) and np.array_equal(dev["yaw"], host["yaw"]), lab
            assert np.array_equal(dev["episode_length_buf"], host["episode_length_buf"]), lab
            assert np.abs(env.completed_terminal_critic_state.cpu().numpy()
                          - host["completed_terminal_critic_state"]).max() <= 2e-5, lab
            assert np.array_equal(env.completed_group_reward.cpu().numpy(), host["completed_group_reward"]), lab
        # the tensor handed out before the steps still holds the pre-step critic state (trainers store it afterwards)
        assert torch.equal(crit, crit_before)
    assert np.abs(env.get_critic_state().cpu().numpy() - oracle.critic_state(p, host)).max() <= 2e-5


def test_fused_critic_equals_standalone_kernel():
    """The critic state written by the step / rollout / reset epilogue is bit-identical to swarm_critic_state's, and
    tensors handed out earlier are not overwritten by later steps."""
    E = 256
    a, b = _mk("shl", "oc2", E, fused_critic=True), _mk("shl", "oc2", E, fused_critic=False)
    rng = np.random.default_rng(5)
    for env in (a, b):
        env.reset(seed=3)
    assert torch.equal(a.get_critic_state(), b.get_critic_state())          # from reset()'s epilogue
    held = []
    for d in range(4):
        act = torch.as_tensor(_actions(rng, a.params, E), device=DEV)
        for t in range(5):
            a.step_tensor(act)
            b.step_tensor(act)
        ca, cb = a.get_critic_state(), b.get_critic_state()
        assert torch.equal(ca, cb), f"decision {d}"
        n0 = _lib.load().swarm_kernel_launch_count()
        assert a.get_critic_state() is ca and _lib.load().swarm_kernel_launch_count() == n0   # same state asked twice
        held.append((ca, cb.clone()))
        for ca_old, cb_old in held[-(a.CRITIC_POOL - 1):]:
            assert torch.equal(ca_old, cb_old)                               # still intact within the pool's horizon
    act = torch.as_tensor(_actions(rng, a.params, E), device=DEV)
    a.rollout(act, 5)
    b.rollout(act, 5)
    assert torch.equal(a.get_critic_state(), b.get_critic_state())          # from the rollout's last step
    # a pose written from outside invalidates the fused copy
    for t in range(5):
        a.step_tensor(act), b.step_tensor(act)
    assert a._critic_fresh                                                   # the fifth step wrote it
    for env in (a, b):
        env.agent_pos[:, 0] += 0.01
    assert torch.equal(a.get_critic_state(), b.get_critic_state())


def test_rebinding_state_attributes_keeps_the_kernel_pointers():
    env = _mk("xor", "cyclamen", 64)
    env.reset(seed=1)
    ptrs = (env.episode_length_buf.data_ptr(), env.agent_pos.data_ptr(), env.agent_yaw.data_ptr())
    env.episode_length_buf = torch.full((64,), env.max_episode_length - 1, dtype=torch.long)
    env.agent_yaw = torch.zeros(64, N)
    assert ptrs == (env.episode_length_buf.data_ptr(), env.agent_pos.data_ptr(), env.agent_yaw.data_ptr())
    _, _, to = env.step_tensor(torch.zeros(64, N, 1, dtype=torch.long, device=DEV))
    assert bool(to.all())                       # the kernel saw the rebound counters: every env timed out
    assert int(env.episode_length_buf.max()) == 0


def test_gym_make_builds_and_steps_the_env():
    """scripts/train.py:188 on the GPU: gym.make(id, cfg=cfg) through the registration (gymnasium's API played by
    tests/gym_stub.py, the package is not in this image) returns a working SwarmEnv."""
    import gym_stub
    from swarmacb_isaaclab_b200 import env as env_mod
    from swarmacb_isaaclab_b200.cfg import TASK_CFGS
    gym = gym_stub.install()
    try:
        assert env_mod.register_gym()
        cfg = TASK_CFGS["SwarmACB-Sheltering-v0"]()
        cfg.update_variant("daisy")
        cfg.scene.num_envs, cfg.sim.device = 32, DEV
        env = gym.make("SwarmACB-SHL-v0", cfg=cfg)          # the alias id of missions/sheltering/__init__.py:8-16
        obs, _ = env.reset()
        act = {a: torch.randint(0, 6, (32, 1), device=DEV) for a in env.possible_agents}
        obs, rew, term, trunc, _ = env.unwrapped.step(act)
        assert obs["epuck_7"].shape == (32, 24) and rew["epuck_0"].shape == (32,)
    finally:
        gym_stub.uninstall()


def test_viewer_feed_reads_one_env_from_device_state():
    from swarmacb_isaaclab_b200 import viewer
    env = _mk("for", "daisy", 128)
    env.reset(seed=2)
    act = torch.randint(0, 6, (128, N, 1), device=DEV)
    for _ in range(3):
        env.step_tensor(act)
    sc, fr = viewer.scene(env), viewer.frame(env, 77)
    assert np.array_equal(fr["pos"], env.agent_pos[77].cpu().numpy()) and fr["episode_step"] == 3
    assert np.array_equal(fr["obs"], env._obs[77].cpu().numpy()) and fr["has_food"].shape == (N,)
    assert viewer.render_svg(sc, fr).count('class="robot"') == N
