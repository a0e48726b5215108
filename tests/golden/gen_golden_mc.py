"""Golden rollouts of scripts/manual_control.py's StandaloneDGTEnv (BASELINE config 1), run in the build container.

    python tests/golden/gen_golden_mc.py        # rewrites tests/golden/mc_*.npz

Replays the per-tick sequence of the reference's pygame loop (manual_control.py:721-757: sensors -> dispatch ->
step -> episode roll-over -> compute_obs_robot0) headless, with scripted module ids / robot-0 wheel commands and
every torch.rand / torch.randint draw recorded, under torch.manual_seed.  XOR runs the full 180 s (1800 ticks).
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refstub  # noqa: E402

N = 20
FSM_FIELDS = ["_explore_state", "_explore_steps", "_explore_dir", "_photo_avoiding", "_photo_steps", "_photo_dir",
              "_antiphoto_avoiding", "_antiphoto_steps", "_antiphoto_dir"]


def load_mc():
    spec = importlib.util.spec_from_file_location("ref_manual_control", refstub.REF_ROOT + "/scripts/manual_control.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def snap(env):
    bm = env.behavior_modules
    s = {"pos": env.pos[0], "yaw": env.yaw[0], "prev_ground": env.prev_ground_color[0],
         "has_food": env.has_food[0], "prev_in_nest": env.prev_in_nest[0]}
    for f in FSM_FIELDS:
        s["fsm" + f] = getattr(bm, f)[0]
    out = {k: v.detach().clone().numpy() for k, v in s.items()}
    out["step_count"] = np.int64(env.step_count)
    out["episode_reward"] = np.float32(env.episode_reward)
    return out


def run(task, ticks, seed, start_count=None):
    mc = load_mc()
    from behavior_modules import BehaviorModules  # the module object manual_control imported
    torch.manual_seed(seed)
    env = mc.StandaloneDGTEnv(N, "cpu", task)
    if start_count is not None:
        env.step_count = start_count
    gen = torch.Generator().manual_seed(seed + 1)
    slots = []
    for name, slot in (("_exploration", 0), ("_phototaxis", 1), ("_anti_phototaxis", 2)):
        if getattr(BehaviorModules, "_mc_tagged", False):
            break
        orig = getattr(BehaviorModules, name)

        def wrapped(self, *a, _orig=orig, _slot=slot, **k):
            n0 = len(TAP[0].randint_calls) if TAP[0] else 0
            out = _orig(self, *a, **k)
            if TAP[0] and len(TAP[0].randint_calls) > n0:
                SLOTS.append(_slot)
            return out

        setattr(BehaviorModules, name, wrapped)
    BehaviorModules._mc_tagged = True

    recs = []
    init = snap(env)
    ids = torch.ones(1, N, dtype=torch.long)
    for t in range(ticks):
        if t % 7 == 0:
            ids = torch.randint(0, 6, (1, N), generator=gen)
        ids[0, 0] = 1
        w0 = (torch.rand(2, generator=gen) * 2 - 1) * 0.2      # beyond +-0.16 sometimes: exercises the clamp
        SLOTS.clear()
        with refstub.NoiseTap() as tap:
            TAP[0] = tap
            left = torch.zeros(1, N)
            right = torch.zeros(1, N)
            left[0, 0], right[0, 0] = w0[0], w0[1]
            _, prox_val, prox_ang = env.sensors.compute_proximity(env.pos, env.yaw, env.wall_segments, env.pos, env.robot_radius)
            _, light_val, light_ang = env._compute_light_readings()
            _, _, rab_ax, rab_ay = env.sensors.compute_rab(env.pos, env.yaw, obstacle_segments=env.wall_segments)
            el, er = env.behavior_modules.dispatch(ids.clone(), prox_val, prox_ang, light_val, light_ang, rab_ax, rab_ay)
            left[0, 1:] = el[0, 1:]
            right[0, 1:] = er[0, 1:]
            env.step(left, right)
            reward = env.step_reward
            rolled = env.step_count >= env.episode_steps
            if rolled:
                env.reset(advance_episode=True)
            info = env.compute_obs_robot0()
            TAP[0] = None
        rab = [r for r in tap.rand_calls if r.dim() == 3]
        spawn = [r for r in tap.rand_calls if r.dim() == 1]
        assert len(rab) == 2 and len(spawn) == (3 if rolled else 0)
        turn_dur = np.ones((N, 3), np.int32)
        assert len(SLOTS) == len(tap.randint_calls)
        for slot, draw in zip(SLOTS, tap.randint_calls):
            turn_dur[:, slot] = draw[0].numpy()
        # full 24-dim observation of all robots for the same second noise draw: recompute deterministically
        rec = {"module_ids": ids[0].numpy().copy(), "wheels0": w0.numpy().copy(),
               "rab_keep1": (rab[0][0] >= 0.85).numpy(), "rab_keep2": (rab[1][0] >= 0.85).numpy(),
               "turn_dur": turn_dur, "reward": np.float32(reward), "rolled": np.bool_(rolled),
               "mc_spawn_u": (np.stack([s.numpy() for s in spawn], axis=1) if rolled else np.zeros((N, 3), np.float32)),
               "obs0": np.asarray(info["obs_24"], np.float32),
               "completed": np.float32(env.completed_episode_reward if env.completed_episode_reward is not None else 0.0)}
        post = snap(env)
        rec.update({"post_" + k: v for k, v in post.items()})
        recs.append(rec)
    arrays = {k: np.stack([r[k] for r in recs]) for k in recs[0]}
    arrays["rab_keep1"] = np.packbits(arrays["rab_keep1"].reshape(ticks, -1), axis=1)
    arrays["rab_keep2"] = np.packbits(arrays["rab_keep2"].reshape(ticks, -1), axis=1)
    for k, v in init.items():
        arrays["init_" + k] = v
    meta = dict(task=task, ticks=ticks, seed=seed, episode_steps=int(env.episode_steps), torch=torch.__version__)
    return meta, arrays


TAP = [None]
SLOTS = []

CASES = [
    ("SwarmACB-XOR-v0", 1800, 500, None),           # BASELINE config 1: the full 180 s rollout
    ("SwarmACB-DirectionalGate-v0", 160, 501, 1120),
    ("SwarmACB-Homing-v0", 160, 502, 1100),
    ("SwarmACB-Foraging-v0", 160, 503, 1700),
    ("SwarmACB-Sheltering-v0", 160, 504, 1720),
]


def main():
    for task, ticks, seed, start in CASES:
        meta, arrays = run(task, ticks, seed, start)
        name = "mc_" + task.split("-")[1].lower() + ".npz"
        path = os.path.join(HERE, name)
        np.savez_compressed(path, meta=np.array(json.dumps(meta)), **arrays)
        print(f"{name:24s} {os.path.getsize(path)/1024:8.1f} KiB  ticks={ticks} rollovers={int(arrays['rolled'].sum())} "
              f"reward_sum={arrays['reward'].sum():.0f}")


if __name__ == "__main__":
    main()
