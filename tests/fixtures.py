"""Loader + comparator for the golden single-step fixtures (tests/golden/*.npz).

A fixture holds ``steps`` teacher-forced reference steps for E envs: state before, actions, recorded
noise, state after, observation, reward, time_out, critic state.  ``step_case`` converts one step to
the ABI layouts of include/swarm_abi.h; ``compare`` applies the parity bar of BASELINE.json
(bit-exact integers/counters, 1e-5 m / 1e-5 rad poses, 1e-4 sensor readings).
"""
from __future__ import annotations

import glob
import json
import os

import numpy as np
import torch

from swarmacb_isaaclab_b200 import MISSION_CFGS, build_params
from swarmacb_isaaclab_b200.params import N, pack_fsm, unpack_fsm

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

POS_TOL = 1e-5
YAW_TOL = 1e-5
SENSOR_TOL = 1e-4
CACHE_FIELDS = ["prox_value", "prox_angle", "light_value", "light_angle", "rab_attr_x", "rab_attr_y"]


def fixture_files():
    return sorted(p for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")) if not os.path.basename(p).startswith("mc_"))


def make_cfg(mission: str, mode: str, num_envs: int, decimation: int = 1, device: str = "cpu"):
    cfg = MISSION_CFGS[mission]()
    if mode in ("oc2", "oc2c"):
        cfg.update_variant("cyclamen")
        cfg.use_continuous_actions(full_observations=(mode == "oc2"))
    else:
        cfg.update_variant(mode)
    cfg.scene.num_envs = num_envs
    cfg.decimation = decimation
    cfg.sim.device = device
    return cfg


class Fixture:
    def __init__(self, path=None, meta=None, arrays=None):
        """From a committed .npz, or (``meta``/``arrays``) straight from tests/golden/gen_golden.make_case-style
        records of the live reference (tests/test_oracle_vs_live_reference.py)."""
        if path is not None:
            z = np.load(path)
            self.path = path
            self.name = os.path.basename(path)[:-4]
            self.meta = json.loads(str(z["meta"]))
            self.z = {k: z[k] for k in z.files if k != "meta"}
        else:
            self.path, self.name, self.meta, self.z = None, "live", dict(meta), dict(arrays)
        m = self.meta
        self.E, self.steps = m["E"], m["steps"]
        self.cfg = make_cfg(m["mission"], m["mode"], m["E"], m["decimation"])
        self.params = build_params(self.cfg)
        assert self.params.obs_dim == m["obs_dim"]
        assert self.params.max_episode_length == m["max_episode_length"]
        self.rab_keep = np.unpackbits(self.z["rab_keep"], axis=1)[:, : self.E * N * N].reshape(
            self.steps, self.E, N, N).astype(bool)

    def _state(self, prefix, t):
        z, E = self.z, self.E
        g = lambda k: z[f"{prefix}_{k}"][t]
        s = {
            "pos": g("pos").astype(np.float32), "yaw": g("yaw").astype(np.float32),
            "prev_ground": g("prev_ground").astype(np.float32),
            "cached_left": g("cached_left").astype(np.float32),
            "cached_right": g("cached_right").astype(np.float32),
            "episode_length_buf": g("ep_len").astype(np.int64),
            "episode_group_reward": g("ep_reward").astype(np.float32),
            "completed_group_reward": g("completed_group_reward").astype(np.float32),
            "completed_terminal_critic_state": g("completed_terminal_critic_state").astype(np.float32),
        }
        fsm = pack_fsm(*[torch.from_numpy(np.ascontiguousarray(g("fsm" + f))) for f in (
            "_explore_state", "_explore_steps", "_explore_dir", "_photo_avoiding", "_photo_steps",
            "_photo_dir", "_antiphoto_avoiding", "_antiphoto_steps", "_antiphoto_dir")])
        s["fsm"] = fsm.numpy().astype(np.int32)
        s["beh_cache"] = np.stack([g("cache_" + f) for f in CACHE_FIELDS], axis=1).astype(np.float32)  # (E,6,N)
        if f"{prefix}_has_food" in z:
            s["mission_flags"] = (g("has_food").astype(np.uint8) | (g("prev_in_nest").astype(np.uint8) << 1))
        else:
            s["mission_flags"] = np.zeros((E, N), np.uint8)
        return {k: np.ascontiguousarray(v) for k, v in s.items()}

    def step_case(self, t):
        z = self.z
        p_loss = float(self.params.rab_loss_probability)
        rab_u = np.where(self.rab_keep[t], np.float32(min(1.0, p_loss + 0.05)), np.float32(p_loss * 0.5))
        actions = z["actions"][t]
        if self.meta["discrete"]:
            actions = actions.reshape(self.E, N).astype(np.int64)
        R = int(z["spawn_rounds"][t])
        return {
            "pre": self._state("pre", t),
            "post": self._state("post", t),
            "actions": actions,
            "rab_u": rab_u.astype(np.float32),
            "turn_dur": z["turn_dur"][t].astype(np.int32),
            "spawn_u": z["spawn_u"][t][: max(R, 1)].astype(np.float32),
            "yaw_u": z["yaw_u"][t].astype(np.float32),
            "obs": z["obs"][t], "reward": z["reward"][t], "time_out": z["time_out"][t].astype(bool),
            "critic_state": z["critic_state"][t],
        }


def _maxdiff(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)))) if np.size(a) else 0.0


def angle_diff(a, b):
    d = np.asarray(a, np.float64) - np.asarray(b, np.float64)
    return np.abs((d + np.pi) % (2 * np.pi) - np.pi)


def compare(case, params, state, obs, reward, time_out, critic=None, label=""):
    """Raise AssertionError with a readable report if the step result violates the parity bar."""
    post = case["post"]
    errs = []

    def check_float(name, got, want, tol):
        d = _maxdiff(got, want)
        if not d <= tol:
            idx = np.unravel_index(np.argmax(np.abs(np.asarray(got, np.float64) - np.asarray(want, np.float64))),
                                   np.shape(want))
            errs.append(f"{name}: max|diff|={d:.3e} > {tol:g} at {idx} got={np.asarray(got)[idx]} want={np.asarray(want)[idx]}")

    def check_exact(name, got, want):
        got, want = np.asarray(got), np.asarray(want)
        if not np.array_equal(got, want):
            bad = np.argwhere(got != want)
            errs.append(f"{name}: {len(bad)} mismatches, first at {tuple(bad[0])} got={got[tuple(bad[0])]} want={want[tuple(bad[0])]}")

    check_float("pos", state["pos"], post["pos"], POS_TOL)
    d = float(angle_diff(state["yaw"], post["yaw"]).max())
    if not d <= YAW_TOL:
        errs.append(f"yaw: max|diff|={d:.3e}")
    check_exact("time_out", np.asarray(time_out).astype(bool), case["time_out"])
    check_exact("reward", reward, case["reward"])
    check_exact("episode_length_buf", state["episode_length_buf"], post["episode_length_buf"])
    check_exact("episode_group_reward", state["episode_group_reward"], post["episode_group_reward"])
    check_exact("completed_group_reward", state["completed_group_reward"], post["completed_group_reward"])
    check_exact("prev_ground", state["prev_ground"], post["prev_ground"])
    check_exact("mission_flags", state["mission_flags"], post["mission_flags"])
    check_float("cached_left", state["cached_left"], post["cached_left"], 1e-6)
    check_float("cached_right", state["cached_right"], post["cached_right"], 1e-6)
    check_float("completed_terminal_critic_state", state["completed_terminal_critic_state"],
                post["completed_terminal_critic_state"], 2e-5)
    if params.discrete_actions:
        got = unpack_fsm(torch.from_numpy(np.ascontiguousarray(state["fsm"])))
        want = unpack_fsm(torch.from_numpy(np.ascontiguousarray(post["fsm"])))
        for k in want:
            check_exact("fsm" + k, got[k].numpy(), want[k].numpy())
        cache_got, cache_want = state["beh_cache"], post["beh_cache"]
        for c, f in enumerate(CACHE_FIELDS):
            if f.endswith("angle"):
                # angles of near-zero vectors are ill-conditioned; compare as vectors instead
                mag_i = c - 1
                vg = cache_got[:, mag_i] * np.stack([np.cos(cache_got[:, c]), np.sin(cache_got[:, c])])
                vw = cache_want[:, mag_i] * np.stack([np.cos(cache_want[:, c]), np.sin(cache_want[:, c])])
                scale = max(1.0, float(np.abs(cache_want[:, mag_i]).max()))
                check_float("cache_" + f + "(vec)", vg, vw, SENSOR_TOL * scale)
            else:
                scale = max(1.0, float(np.abs(cache_want[:, c]).max())) if f == "light_value" else 1.0
                check_float("cache_" + f, cache_got[:, c], cache_want[:, c], SENSOR_TOL * scale)
    check_float("obs", obs, case["obs"], SENSOR_TOL)
    if critic is not None:
        check_float("critic_state", critic, case["critic_state"], 2e-5)
    if errs:
        raise AssertionError(f"[{label}] parity violations:\n  " + "\n  ".join(errs))


# ── manual_control.py (BASELINE config 1) rollouts ─────────────────────────────────────────────
MC_FSM = ["_explore_state", "_explore_steps", "_explore_dir", "_photo_avoiding", "_photo_steps", "_photo_dir",
          "_antiphoto_avoiding", "_antiphoto_steps", "_antiphoto_dir"]


def mc_fixture_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "mc_*.npz")))


class McFixture:
    """Tick-by-tick record of the reference's StandaloneDGTEnv loop (tests/golden/gen_golden_mc.py), E = 1."""

    def __init__(self, path):
        from swarmacb_isaaclab_b200.params import build_mc_params
        z = np.load(path)
        self.name = os.path.basename(path)[:-4]
        self.meta = json.loads(str(z["meta"]))
        self.z = {k: z[k] for k in z.files if k != "meta"}
        self.ticks = self.meta["ticks"]
        self.params = build_mc_params(self.meta["task"])
        assert self.params.max_episode_length == self.meta["episode_steps"]
        unpack = lambda a: np.unpackbits(a, axis=1)[:, : N * N].reshape(self.ticks, 1, N, N).astype(bool)
        self.keep1, self.keep2 = unpack(self.z["rab_keep1"]), unpack(self.z["rab_keep2"])

    def state(self, t):
        """State BEFORE tick t (ABI layouts, E=1)."""
        z = self.z
        g = (lambda k: z["init_" + k]) if t == 0 else (lambda k: z["post_" + k][t - 1])
        fsm = pack_fsm(*[torch.from_numpy(np.ascontiguousarray(g("fsm" + f))) for f in MC_FSM]).numpy().astype(np.int32)
        return {
            "pos": g("pos").astype(np.float32)[None], "yaw": g("yaw").astype(np.float32)[None],
            "prev_ground": g("prev_ground").astype(np.float32)[None],
            "cached_left": np.zeros((1, N), np.float32), "cached_right": np.zeros((1, N), np.float32),
            "fsm": fsm[None], "beh_cache": np.zeros((1, 6, N), np.float32),
            "mission_flags": (g("has_food").astype(np.uint8) | (g("prev_in_nest").astype(np.uint8) << 1))[None],
            "episode_length_buf": np.array([g("step_count")], np.int64),
            "episode_group_reward": np.array([g("episode_reward")], np.float32),
            "completed_group_reward": np.zeros(1, np.float32),
            "completed_terminal_critic_state": np.zeros((1, N, 5), np.float32),
        }

    def tick_inputs(self, t):
        z = self.z
        p_loss = float(self.params.rab_loss_probability)
        u = lambda keep: np.where(keep, np.float32(min(1.0, p_loss + 0.05)), np.float32(p_loss * 0.5)).astype(np.float32)
        wheels = np.zeros((1, N, 2), np.float32)
        wheels[0, 0] = z["wheels0"][t]
        return dict(module_ids=z["module_ids"][t][None].astype(np.int64), wheels=wheels, rab_u=u(self.keep1[t]),
                    rab_u2=u(self.keep2[t]), turn_dur=z["turn_dur"][t][None].astype(np.int32),
                    mc_spawn_u=z["mc_spawn_u"][t][None].astype(np.float32))

    def check(self, t, state, obs, reward, rolled, label=""):
        z, want = self.z, self.state(t + 1)
        errs = []
        d = np.abs(state["pos"] - want["pos"]).max()
        if not d <= POS_TOL:
            errs.append(f"pos {d:.3e}")
        d = angle_diff(state["yaw"], want["yaw"]).max()
        if not d <= YAW_TOL:
            errs.append(f"yaw {d:.3e}")
        keys = ["fsm", "prev_ground", "episode_length_buf", "episode_group_reward"]
        if self.params.mission == 3:  # has_food / prev_in_nest only mean something in Foraging (MC:385-392)
            keys.append("mission_flags")
        for k in keys:
            if not np.array_equal(state[k], want[k]):
                errs.append(f"{k} mismatch: got {state[k].ravel()[:6]} want {want[k].ravel()[:6]}")
        if float(reward[0]) != float(z["reward"][t]):
            errs.append(f"reward {reward[0]} != {z['reward'][t]}")
        if bool(rolled[0]) != bool(z["rolled"][t]):
            errs.append("episode roll-over flag")
        d = np.abs(obs[0, 0] - z["obs0"][t]).max()
        if not d <= SENSOR_TOL:
            errs.append(f"obs robot0 {d:.3e}")
        if errs:
            raise AssertionError(f"[{label} tick {t}] " + "; ".join(errs))
