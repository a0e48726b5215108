"""The reference's own trainers, unmodified, on top of ``SwarmEnv``'s host logic (BASELINE config 5 at protocol level).

``scripts/train.py:109-207`` is replayed headless through ``swarmacb_isaaclab_b200.runner``: YAML -> env cfg
-> env -> ``POCATrainer`` / ``FixedOptionCriticTrainer`` / ``LearnedOptionCriticTrainer`` imported from
``/root/reference``.  The build container has no GPU, so the C-ABI call sites are served by the CPU oracle
(tests/oracle_env.py); what is under test is everything the trainers touch: the dict API, ``unwrapped.*``
attributes, action views, counters, auto-reset and ``completed_*`` buffers.  Skipped where the reference tree is
absent (the GPU box).
"""
import os

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


def _shrink(cfg):
    cfg.horizon = 6
    cfg.mini_batch_size = 64
    cfg.num_epochs = 1
    cfg.summary_freq = 10 ** 9
    cfg.checkpoint_interval = 10 ** 9
    cfg.hidden_dim = 32
    cfg.critic_hidden_dim = 32
    if hasattr(cfg, "option_hidden_dim"):
        cfg.option_hidden_dim = 32
    if hasattr(cfg, "sequence_length"):
        cfg.sequence_length = 4
    if hasattr(cfg, "fused_optimizer"):
        cfg.fused_optimizer = False
    cfg.buffer_size_hint = 0


@pytest.mark.parametrize("yaml_name", [
    "XOR_cyclamen.yaml",             # MA-POCA, recurrent discrete actor, 4-dim obs
    "Foraging_dandelion.yaml",       # MA-POCA, continuous wheels, 24-dim obs
    "OC_Homing_cyclamen.yaml",       # fixed Option-Critic over the six modules
    "OC2_Sheltering_cyclamen.yaml",  # learned Option-Critic: continuous primitives + full observations
])
def test_reference_trainer_runs_unmodified(yaml_name, tmp_path):
    from oracle_env import OracleBackedEnv
    from swarmacb_isaaclab_b200 import runner
    from swarmacb_isaaclab_b200.params import N

    path = os.path.join(REF, "configs", yaml_name)
    spec = runner.load_run_spec(path)
    agents = runner.import_reference_agents(REF)
    # our spec parser agrees with the reference's loader on the env-facing fields
    run_name, variant, ref_cfg, env_ov = agents.load_config(path)
    assert (spec.run_name, spec.variant, spec.trainer_type) == (run_name, variant, ref_cfg.trainer_type)
    assert spec.task_id == env_ov.get("task") and spec.decision_period == ref_cfg.decision_period
    assert spec.env_overrides == {k: v for k, v in env_ov.items() if k != "task"}

    E = 3
    spec.env_overrides["num_envs"] = E
    spec.env_overrides["episode_length_s"] = 2.0       # 20 motion updates: several auto-resets inside the run
    cfg = runner.build_env_cfg(spec.task_id, spec.variant, spec.trainer_type, spec.env_overrides, seed=0, device="cpu")
    env = OracleBackedEnv(cfg)
    assert env.max_episode_length == 20
    torch.manual_seed(0)
    decisions = 14
    trainer = runner.run_reference_trainer(env, agents, path, total_timesteps=E * N * decisions, seed=0,
                                           log_dir=str(tmp_path / "runs"), checkpoint_dir=str(tmp_path / "ckpt"),
                                           tweak=_shrink)
    lib = env._lib
    assert lib.steps >= decisions * spec.decision_period          # every decision held for decision_period steps
    assert int(env.episode_length_buf.max()) < env.max_episode_length
    assert torch.isfinite(env._obs).all() and env._obs.shape == (E, N, env.obs_dim)
    assert float(env.completed_terminal_critic_state.abs().sum()) > 0   # time-outs left the pre-reset critic state
    assert any((tmp_path / "ckpt").iterdir()), "trainer wrote no checkpoint"
    assert trainer is not None
