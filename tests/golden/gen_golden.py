"""Generate golden single-step fixtures from the UNMODIFIED reference (run in the build container).

    python tests/golden/gen_golden.py            # rewrites tests/golden/*.npz

The reference env classes (imported from /root/reference through ``refstub``) are stepped on
CPU with every stochastic draw recorded (``torch.rand`` / ``torch.randint``), and each step is
stored as (state before, actions, noise, state after, obs, reward, time_out, critic state) so the
oracle and the CUDA kernel can be checked teacher-forced, one step at a time
(SURVEY.md §8c).  The committed ``.npz`` files travel to the GPU box; this script and
``refstub`` do not run there.
"""
from __future__ import annotations

import json
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refstub  # noqa: E402

N = 20
FSM_FIELDS = [
    "_explore_state", "_explore_steps", "_explore_dir",
    "_photo_avoiding", "_photo_steps", "_photo_dir",
    "_antiphoto_avoiding", "_antiphoto_steps", "_antiphoto_dir",
]
CACHE_FIELDS = ["prox_value", "prox_angle", "light_value", "light_angle", "rab_attr_x", "rab_attr_y"]


def snapshot(env):
    bm = env.behavior_modules
    s = {
        "pos": env.agent_pos, "yaw": env.agent_yaw, "prev_ground": env.prev_ground_color,
        "cached_left": env._cached_left_vel, "cached_right": env._cached_right_vel,
        "ep_len": env.episode_length_buf, "ep_reward": env._episode_group_reward,
        "completed_group_reward": env.completed_group_reward,
        "completed_terminal_critic_state": env.completed_terminal_critic_state,
    }
    for f in FSM_FIELDS:
        s["fsm" + f] = getattr(bm, f)
    cache = env._sensor_cache
    assert cache is not None
    for f in CACHE_FIELDS:
        s["cache_" + f] = cache[f]
    if hasattr(env, "_has_food"):
        s["has_food"] = env._has_food
        s["prev_in_nest"] = env._prev_in_nest
    return {k: v.detach().clone().numpy() for k, v in s.items()}


def install_tags(tap_holder):
    """Tag randint draws with the behaviour module that asked for them (instrumentation only)."""
    from SwarmACB_isaac.tasks.direct.epuck.behavior_modules import BehaviorModules
    if getattr(BehaviorModules, "_tagged", False):
        return
    for name, slot in (("_exploration", 0), ("_phototaxis", 1), ("_anti_phototaxis", 2)):
        orig = getattr(BehaviorModules, name)

        def wrapped(self, *a, _orig=orig, _slot=slot, **k):
            tap_holder["slot"] = _slot
            n0 = len(tap_holder["tap"].randint_calls) if tap_holder["tap"] else 0
            out = _orig(self, *a, **k)
            if tap_holder["tap"] and len(tap_holder["tap"].randint_calls) > n0:
                tap_holder["slots"].append(_slot)
            return out

        setattr(BehaviorModules, name, wrapped)
    BehaviorModules._tagged = True


TAP = {"tap": None, "slot": -1, "slots": []}


def record_step(env, actions):
    """Run one reference step; return dict of arrays for this step."""
    E = env.num_envs
    cfg = env.cfg
    pre = snapshot(env)
    act_dict = {a: actions[:, i] for i, a in enumerate(cfg.possible_agents)}
    TAP["slots"] = []
    with refstub.NoiseTap() as tap:
        TAP["tap"] = tap
        obs, rew, term, trunc, _ = env.step(act_dict)
        TAP["tap"] = None
    post = snapshot(env)
    a0 = cfg.possible_agents[0]
    time_out = trunc[a0].numpy().copy()

    turn_dur = np.ones((E, N, 3), dtype=np.int32)
    assert len(tap.randint_calls) == len(TAP["slots"])
    for slot, draw in zip(TAP["slots"], tap.randint_calls):
        turn_dur[:, :, slot] = draw.numpy()

    rab = [r for r in tap.rand_calls if r.dim() == 3 and r.shape[-1] == N and r.shape[-2] == N]
    spawn = [r for r in tap.rand_calls if r.dim() == 3 and r.shape[-1] == 2]
    yaw = [r for r in tap.rand_calls if r.dim() == 2]
    assert len(rab) == 1 and len(rab) + len(spawn) + len(yaw) == len(tap.rand_calls)
    rab_keep = (rab[0] >= cfg.rab_loss_probability).numpy()
    reset_ids = np.nonzero(time_out)[0]
    rounds = len(spawn)
    spawn_u = np.zeros((max(rounds, 1), E, N, 2), dtype=np.float32)
    yaw_u = np.zeros((E, N), dtype=np.float32)
    if rounds:
        assert len(yaw) == 1
        for r, s in enumerate(spawn):
            spawn_u[r, reset_ids] = s.numpy()
        yaw_u[reset_ids] = yaw[0].numpy()

    out = {"pre_" + k: v for k, v in pre.items()}
    out.update({"post_" + k: v for k, v in post.items()})
    out["actions"] = actions.numpy().copy()
    out["turn_dur"] = turn_dur
    out["rab_keep"] = rab_keep
    out["spawn_u"] = spawn_u
    out["spawn_rounds"] = np.int32(rounds)
    out["yaw_u"] = yaw_u
    out["obs"] = torch.stack([obs[a] for a in cfg.possible_agents], dim=1).numpy().copy()
    out["reward"] = rew[a0].numpy().copy()
    out["time_out"] = time_out
    out["critic_state"] = env.get_critic_state().numpy().copy()
    for k in ("prox_vals", "light_vals", "ztilde", "rab_proj"):
        out["sens_" + k] = env._sensor_cache[k].numpy().copy()
    return out


def random_actions(env, gen):
    E = env.num_envs
    if env.cfg.discrete_actions:
        return torch.randint(0, 6, (E, N, 1), generator=gen)
    # slightly beyond [-1, 1] so the clamp (directional_gate_env.py:807) is exercised
    return torch.rand(E, N, 2, generator=gen) * 2.4 - 1.2


CLUSTER_CENTRES = {
    "dgt": [(-0.25, 0.10), (0.22, -0.05), (0.0, 0.30), (1.10, 0.40)],
    "xor": [(-0.50, 0.0), (0.35, 0.20), (-0.24, 0.10), (0.0, -1.12)],
    "hom": [(0.0, -0.70), (0.0, -0.42), (1.13, 0.35), (-0.8, 0.8)],
    "for": [(-0.75, 0.0), (0.70, -0.12), (0.0, -0.60), (0.85, -0.85)],
    "shl": [(0.0, 0.0), (-0.27, 0.05), (0.20, 0.18), (0.0, -0.17)],
}


def scramble(env, mission, gen, radius=0.17):
    """Teacher-forced crowded state: clustered robots + randomised FSM / wheel / colour state."""
    E = env.num_envs
    centres = CLUSTER_CENTRES[mission]
    for e in range(E):
        c = torch.tensor(centres[e % len(centres)])
        r = radius * torch.sqrt(torch.rand(N, generator=gen))
        th = torch.rand(N, generator=gen) * 2 * math.pi
        env.agent_pos[e, :, 0] = c[0] + r * torch.cos(th)
        env.agent_pos[e, :, 1] = c[1] + r * torch.sin(th)
    env.agent_yaw[:] = torch.rand(E, N, generator=gen) * 2 * math.pi - math.pi
    bm = env.behavior_modules
    bm._explore_state = torch.randint(0, 2, (E, N), generator=gen)
    bm._explore_steps = torch.where(bm._explore_state == 1, torch.randint(1, 5, (E, N), generator=gen), 0)
    bm._explore_dir = torch.randint(0, 2, (E, N), generator=gen).float() * 2 - 1
    bm._photo_avoiding = torch.randint(0, 2, (E, N), generator=gen).bool()
    bm._photo_steps = torch.where(bm._photo_avoiding, torch.randint(1, 5, (E, N), generator=gen), 0)
    bm._photo_dir = torch.randint(0, 2, (E, N), generator=gen).float() * 2 - 1
    bm._antiphoto_avoiding = torch.randint(0, 2, (E, N), generator=gen).bool()
    bm._antiphoto_steps = torch.where(bm._antiphoto_avoiding, torch.randint(1, 5, (E, N), generator=gen), 0)
    bm._antiphoto_dir = torch.randint(0, 2, (E, N), generator=gen).float() * 2 - 1
    env._cached_left_vel = (torch.rand(E, N, generator=gen) * 2 - 1) * 0.16
    env._cached_right_vel = (torch.rand(E, N, generator=gen) * 2 - 1) * 0.16
    env.prev_ground_color = torch.randint(0, 3, (E, N), generator=gen).float() * 0.5
    if hasattr(env, "_has_food"):
        env._has_food = torch.randint(0, 2, (E, N), generator=gen).bool()
        env._prev_in_nest = torch.randint(0, 2, (E, N), generator=gen).bool()
    env._episode_group_reward = torch.randint(0, 50, (E,), generator=gen).float()
    env.episode_length_buf[:] = torch.randint(3, 500, (E,), generator=gen)
    env._sensor_cache = None
    env._get_observations()


def make_case(mission, mode, scenario, E, steps, seed, decimation=1):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    env = refstub.make_ref_env(mission, mode, E, decimation=decimation)
    install_tags(TAP)
    env.reset()
    if scenario == "fresh":
        pass
    elif scenario == "crowded":
        scramble(env, mission, gen)
    elif scenario == "timeout":
        for _ in range(3):  # leave the spawn state behind
            env.step({a: random_actions(env, gen)[:, i] for i, a in enumerate(env.cfg.possible_agents)})
        L = env.max_episode_length
        lens = [L - 2, L - 1, 7, L - 2, L - 1, L - 3]
        env.episode_length_buf[:] = torch.tensor([lens[e % len(lens)] for e in range(E)])
    elif scenario == "alltimeout":
        scramble(env, mission, gen, radius=0.3)
        env.episode_length_buf[:] = env.max_episode_length - 2
    else:
        raise ValueError(scenario)
    recs = [record_step(env, random_actions(env, gen)) for _ in range(steps)]
    cfg = env.cfg
    meta = dict(mission=mission, mode=mode, scenario=scenario, E=E, N=N, steps=steps, seed=seed,
                decimation=decimation, discrete=bool(cfg.discrete_actions),
                obs_dim=int(recs[0]["obs"].shape[-1]), max_episode_length=int(env.max_episode_length),
                torch=torch.__version__)
    max_rounds = max(r["spawn_u"].shape[0] for r in recs)
    for r in recs:
        pad = max_rounds - r["spawn_u"].shape[0]
        if pad:
            r["spawn_u"] = np.concatenate([r["spawn_u"], np.zeros((pad,) + r["spawn_u"].shape[1:], np.float32)])
    arrays = {k: np.stack([r[k] for r in recs]) for k in recs[0]}
    arrays["rab_keep"] = np.packbits(arrays["rab_keep"].reshape(steps, -1), axis=1)
    return meta, arrays


def case_list():
    cases = []
    seed = 100
    for mission in ("dgt", "xor", "hom", "for", "shl"):
        for mode in ("cyclamen", "daisy", "dandelion", "oc2", "oc2c"):
            cases.append((mission, mode, "fresh", 3, 2, 1))
        for mode in ("cyclamen", "daisy", "dandelion", "oc2"):
            cases.append((mission, mode, "crowded", 4, 3, 1))
        for mode in ("lily", "dandelion"):
            cases.append((mission, mode, "timeout", 6, 3, 1))
        cases.append((mission, "daisy", "alltimeout", 3, 3, 1))
        cases.append((mission, "daisy", "crowded", 3, 2, 6))
        cases.append((mission, "oc2", "crowded", 3, 2, 3))
    out = []
    for c in cases:
        seed += 1
        out.append(c + (seed,))
    return out


def main():
    total = 0
    for mission, mode, scenario, E, steps, dec, seed in case_list():
        meta, arrays = make_case(mission, mode, scenario, E, steps, seed, decimation=dec)
        name = f"{mission}_{mode}_{scenario}_d{dec}.npz"
        path = os.path.join(HERE, name)
        np.savez_compressed(path, meta=np.array(json.dumps(meta)), **arrays)
        sz = os.path.getsize(path)
        total += sz
        print(f"{name:42s} {sz/1024:7.1f} KiB  resets={int(arrays['time_out'].sum())} "
              f"rounds={int(arrays['spawn_rounds'].max())} rew={arrays['reward'].sum():.0f}")
    print(f"total {total/1e6:.2f} MB")


if __name__ == "__main__":
    main()
