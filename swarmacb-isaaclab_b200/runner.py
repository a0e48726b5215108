"""Headless runner: YAML -> env cfg -> ``SwarmEnv`` -> trainer, without Omniverse.

What ``scripts/train.py:109-207`` of the reference does between ``load_config`` and ``trainer.train()``,
minus ``AppLauncher``: the behaviour block of an ML-Agents-style YAML names the task, the CASA variant, the
trainer type and the ``environment:`` overrides; the env cfg is built exactly as ``scripts/train.py:166-185``
builds it and the env comes from :func:`swarmacb_isaaclab_b200.env.make`.

The trainers themselves (MA-POCA, Option-Critic, learned Option-Critic) are the reference's own PyTorch
code and are NOT part of this package: with ``--reference-root`` pointing at a checkout of the reference they
are imported from there, unmodified, and run on top of the fused env; without it the runner drives the env
with a random policy at the trainers' cadence (one decision every ``decision_period`` motion updates) and
reports agent-steps/s.

    python -m swarmacb_isaaclab_b200.runner --config OC2_Sheltering_cyclamen.yaml --num-envs 16384 \
        [--reference-root /path/to/SwarmACB-isaaclab] [--total-timesteps N] [--decisions N]
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import time
import types
from dataclasses import dataclass, field
from typing import Any

from .cfg import TASK_CFGS

LEARNED_OC_ALIASES = ("learned_option_critic", "option_critic_2", "learned_oc", "oc2")
FIXED_OC_ALIASES = ("option_critic", "fixed_option_critic", "fixed_oc", "oc")
DEFAULT_TASK = "SwarmACB-DirectionalGate-v0"   # scripts/train.py:153


@dataclass
class RunSpec:
    """The env-facing part of a training YAML (agents/config_loader.py:30-187 of the reference)."""
    run_name: str
    task_id: str
    variant: str
    trainer_type: str
    decision_period: int = 5
    env_overrides: dict = field(default_factory=dict)
    max_steps: int | None = None
    time_horizon: int | None = None


def load_run_spec(path: str) -> RunSpec:
    """Parse the first behaviour block of an ML-Agents-style YAML the way ``load_config`` does: ``task`` may sit
    in the block or under ``environment:``; every other ``environment:`` key except ``decision_period`` becomes
    an env-cfg override."""
    import yaml
    if not os.path.exists(path):
        raise FileNotFoundError(f"Config file not found: {path}")
    with open(path, "r", encoding="utf-8") as f:
        raw = yaml.safe_load(f)
    behaviors = raw.get("behaviors", raw)
    if not behaviors:
        raise ValueError("Config must have a top-level 'behaviors' key.")
    run_name = next(iter(behaviors))
    block = behaviors[run_name]
    environment = block.get("environment", {}) or {}
    tt = str(block.get("trainer_type", "poca")).lower()
    if tt in LEARNED_OC_ALIASES:
        tt = "learned_option_critic"
    elif tt in FIXED_OC_ALIASES:
        tt = "option_critic"
    elif tt != "poca":
        raise ValueError(f"Unsupported trainer_type: {tt}")
    overrides = {k: v for k, v in environment.items() if k not in ("task", "decision_period")}
    return RunSpec(run_name=run_name, task_id=block.get("task", environment.get("task")) or DEFAULT_TASK,
                   variant=block.get("variant", "dandelion"), trainer_type=tt,
                   decision_period=int(environment.get("decision_period", 5)), env_overrides=overrides,
                   max_steps=block.get("max_steps"), time_horizon=block.get("time_horizon"))


def build_env_cfg(task_id: str, variant: str, trainer_type: str = "poca", env_overrides: dict | None = None,
                  seed: int = 0, device: str = "cuda:0", warn=print):
    """scripts/train.py:166-185: instantiate the task's cfg, ``update_variant``, continuous primitive actions +
    full observations for the learned Option-Critic, then the YAML overrides (``num_envs`` goes to
    ``scene.num_envs``; unknown keys are reported and ignored)."""
    if task_id not in TASK_CFGS:
        raise KeyError(f"unknown task {task_id!r}; known: {sorted(TASK_CFGS)}")
    cfg = TASK_CFGS[task_id]()
    cfg.seed = seed
    cfg.sim.device = device
    cfg.update_variant(variant)
    if trainer_type == "learned_option_critic":
        cfg.use_continuous_actions(full_observations=True)
    for key, value in (env_overrides or {}).items():
        if key == "num_envs":
            cfg.scene.num_envs = int(value)
        elif hasattr(cfg, key):
            setattr(cfg, key, value)
        else:
            warn(f"[runner] Warning: ignored unknown environment override {key!r}")
    return cfg


def import_reference_agents(reference_root: str):
    """Import the reference's ``tasks/direct/agents`` package from a checkout, unmodified.  Its parents'
    ``__init__`` files need Omniverse, so they are replaced by bare namespace packages (the agents package
    itself imports only torch, yaml, tqdm and tensorboard)."""
    # a git checkout (source/SwarmACB_isaac/SwarmACB_isaac) or a `pip install --target` directory (SwarmACB_isaac)
    for pkg in (os.path.join(reference_root, "source", "SwarmACB_isaac", "SwarmACB_isaac"),
                os.path.join(reference_root, "SwarmACB_isaac")):
        if os.path.isdir(os.path.join(pkg, "tasks", "direct", "agents")):
            break
    else:
        raise FileNotFoundError(f"no SwarmACB_isaac/tasks/direct/agents under {reference_root}")
    for name, path in (("SwarmACB_isaac", pkg), ("SwarmACB_isaac.tasks", os.path.join(pkg, "tasks")),
                       ("SwarmACB_isaac.tasks.direct", os.path.join(pkg, "tasks", "direct"))):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [path]
            sys.modules[name] = m
    return importlib.import_module("SwarmACB_isaac.tasks.direct.agents")


def capture_rollout_stats(trainer, stats: dict):
    """Record what the (unmodified) trainer measures about its own rollouts: the learned Option-Critic's
    ``rollout_sps`` / ``rollout_seconds`` / ``update_seconds`` (agents/learned_option_critic_trainer.py:1840-1855,
    handed to ``_write_update_diagnostics``), and for every trainer the wall clock of each ``collect_rollout`` call
    with the device drained on both sides."""
    import torch
    stats.setdefault("rollout_sps", [])
    stats.setdefault("rollout_seconds", [])
    stats.setdefault("update_seconds", [])
    stats.setdefault("collect_rollout_calls", [])
    inner = trainer.collect_rollout

    def timed_collect(*a, **k):
        torch.cuda.synchronize()
        t0, g0 = time.perf_counter(), trainer.global_step
        out = inner(*a, **k)
        torch.cuda.synchronize()
        stats["collect_rollout_calls"].append({"seconds": time.perf_counter() - t0,
                                               "agent_decisions": int(trainer.global_step - g0)})
        return out

    trainer.collect_rollout = timed_collect
    if hasattr(trainer, "_write_update_diagnostics"):
        diag = trainer._write_update_diagnostics

        def tapped(metrics):
            for k in ("rollout_sps", "rollout_seconds", "update_seconds"):
                if k in metrics:
                    stats[k].append(float(metrics[k]))
            return diag(metrics)

        trainer._write_update_diagnostics = tapped


def run_reference_trainer(env, agents, yaml_path: str, *, total_timesteps: int | None = None, seed: int | None = None,
                          log_dir: str | None = None, checkpoint_dir: str | None = None, tweak=None, hook=None):
    """scripts/train.py:191-203 with the reference's own loader and trainer classes."""
    _, _, cfg, _ = agents.load_config(yaml_path)
    if total_timesteps is not None:
        cfg.total_timesteps = int(total_timesteps)
    if seed is not None:
        cfg.seed = int(seed)
    if log_dir is not None:
        cfg.log_dir = log_dir
    if checkpoint_dir is not None:
        cfg.checkpoint_dir = checkpoint_dir
    if tweak is not None:
        tweak(cfg)
    trainer_type = getattr(cfg, "trainer_type", "poca")
    cls = {"learned_option_critic": "LearnedOptionCriticTrainer", "option_critic": "FixedOptionCriticTrainer",
           "poca": "POCATrainer"}[trainer_type]
    trainer = getattr(agents, cls)(env, cfg)
    if hook is not None:
        hook(trainer)
    trainer.train()
    return trainer


def random_policy_rollout(env, decisions: int, decision_period: int, seed: int = 1) -> dict:
    """Drive the env at the trainers' cadence (agents/poca_trainer.py:551-573): one random action per decision,
    held for ``decision_period`` motion updates via ``SwarmEnv.rollout``; the decision's observation and critic
    state are fetched like the trainers do.  Returns throughput and the episode metrics."""
    import torch
    from .params import N
    E, dev = env.num_envs, env.device
    g = torch.Generator(device=dev).manual_seed(seed)
    discrete = bool(env.params.discrete_actions)
    env.reset()
    total_reward = torch.zeros((), dtype=torch.float64, device=dev)
    episodes = torch.zeros((), dtype=torch.int64, device=dev)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(decisions):
        if discrete:
            act = torch.randint(0, 6, (E, N, 1), generator=g, device=dev)
        else:
            act = torch.rand(E, N, 2, generator=g, device=dev) * 2 - 1
        env.get_critic_state()
        _, reward, time_out = env.rollout(act, decision_period)
        total_reward += reward.sum()
        episodes += time_out.sum()
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    steps = decisions * decision_period
    return {"envs": E, "decisions": decisions, "decision_period": decision_period, "env_steps": steps,
            "agent_steps_per_s": E * N * steps / dt, "agent_decisions_per_s": E * N * decisions / dt,
            "seconds": dt, "sum_group_reward": float(total_reward), "episodes_finished": int(episodes)}


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config", required=True, help="ML-Agents-style YAML (the reference's configs/*.yaml format)")
    ap.add_argument("--task", default=None)
    ap.add_argument("--variant", default=None)
    ap.add_argument("--num-envs", type=int, default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--reference-root", default=None, help="checkout of the reference: run ITS trainer on this env")
    ap.add_argument("--total-timesteps", type=int, default=None)
    ap.add_argument("--log-dir", default=None)
    ap.add_argument("--checkpoint-dir", default=None)
    ap.add_argument("--decisions", type=int, default=200, help="random-policy mode: decisions to run")
    ap.add_argument("--horizon", type=int, default=None,
                    help="trainer mode: override time_horizon (SURVEY 7.8: the OC2 buffer needs <= ~40 at 16384 envs)")
    ap.add_argument("--updates", type=int, default=None,
                    help="trainer mode: stop after this many rollout+update iterations (sets total_timesteps)")
    ap.add_argument("--epochs", type=int, default=None, help="trainer mode: override num_epochs")
    ap.add_argument("--report-json", default=None, help="trainer mode: write the captured rollout statistics here")
    args = ap.parse_args(argv)

    spec = load_run_spec(args.config)
    if args.variant:
        spec.variant = args.variant
    if args.task:
        spec.task_id = args.task
    if args.num_envs is not None:
        spec.env_overrides["num_envs"] = args.num_envs
    cfg = build_env_cfg(spec.task_id, spec.variant, spec.trainer_type, spec.env_overrides, args.seed, args.device)
    from .env import make
    env = make(spec.task_id, cfg=cfg)
    if args.reference_root:
        from .params import N
        agents = import_reference_agents(args.reference_root)
        stats: dict[str, Any] = {}

        def tweak(tcfg):
            if args.horizon is not None:
                tcfg.horizon = int(args.horizon)
                if hasattr(tcfg, "sequence_length"):
                    tcfg.sequence_length = min(int(tcfg.sequence_length), int(args.horizon))
            if args.epochs is not None:
                tcfg.num_epochs = int(args.epochs)
            if args.updates is not None:
                # one iteration collects `horizon` decisions of every agent, then updates once
                tcfg.buffer_size_hint = 0
                tcfg.total_timesteps = int(args.updates) * int(tcfg.horizon) * env.num_envs * N
            tcfg.summary_freq = max(int(getattr(tcfg, "summary_freq", 1)), 1)

        t0 = time.perf_counter()
        trainer = run_reference_trainer(env, agents, args.config, total_timesteps=args.total_timesteps, seed=args.seed,
                                        log_dir=args.log_dir, checkpoint_dir=args.checkpoint_dir, tweak=tweak,
                                        hook=lambda tr: capture_rollout_stats(tr, stats))
        period = int(getattr(trainer, "decision_period", spec.decision_period))
        calls = stats.get("collect_rollout_calls", [])
        secs = sum(c["seconds"] for c in calls)
        decs = sum(c["agent_decisions"] for c in calls)
        report = {"run": spec.run_name, "task": spec.task_id, "variant": spec.variant, "trainer_type": spec.trainer_type,
                  "trainer_class": type(trainer).__name__, "envs": env.num_envs, "decision_period": period,
                  "horizon": int(trainer.cfg.horizon), "global_step": int(trainer.global_step),
                  "wall_seconds": time.perf_counter() - t0,
                  "trainer_rollout_sps": stats.get("rollout_sps"), "trainer_rollout_seconds": stats.get("rollout_seconds"),
                  "trainer_update_seconds": stats.get("update_seconds"),
                  "collect_rollout": {"calls": len(calls), "seconds": secs, "agent_decisions": decs,
                                      "agent_decisions_per_s": decs / secs if secs else None,
                                      "agent_steps_per_s": decs * period / secs if secs else None}}
        print(json.dumps(report))
        if args.report_json:
            with open(args.report_json, "w") as f:
                json.dump(report, f, indent=1)
    else:
        out: dict[str, Any] = {"run": spec.run_name, "task": spec.task_id, "variant": spec.variant,
                               "trainer_type": spec.trainer_type, "obs_dim": env.obs_dim,
                               "discrete_actions": bool(env.params.discrete_actions)}
        out.update(random_policy_rollout(env, args.decisions, spec.decision_period, seed=args.seed + 1))
        print(json.dumps(out))
    env.close()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
