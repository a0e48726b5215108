"""ctypes mirrors of include/swarm_abi.h and the host-side derivation of ``SwarmParams``.

Every constant is derived the way the reference derives it - Python double arithmetic, rounded to
float32 at the point where torch would cast the Python scalar (ctypes ``c_float`` assignment rounds
to nearest) - so threshold tests in the kernel see the same float32 values as the reference's
tensors.  Citations: ENV = directional_gate_env.py, SENS = epuck_sensors.py, CFG =
directional_gate_env_cfg.py of the reference.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

ABI_VERSION = 2
FSM_STATE_MASK = (1 << 18) - 1   # fsm word: bits 0..17 reference state, 18..23 pre-drawn turn-duration bits
N = 20
MAX_SEG = 16
MAX_INTERNAL = 4

MISSION_ID = {"dgt": 0, "xor": 1, "hom": 2, "for": 3, "shl": 4}
GATE_NONE, GATE_DGT, GATE_SHL = 0, 1, 2

F = C.c_float
I = C.c_int32


class SwarmParams(C.Structure):
    _fields_ = [
        ("abi_version", I), ("mission", I), ("obs_dim", I), ("discrete_actions", I),
        ("decimation", I), ("max_episode_length", I), ("solver_iterations", I), ("has_light", I),
        ("n_segments", I), ("n_internal", I), ("gate_mode", I), ("spawn_max_attempts", I),
        ("dt", F), ("wheelbase", F), ("max_wheel_speed", F),
        ("robot_radius", F), ("robot_radius_sq", F), ("two_radius", F),
        ("wall_r_eff", F), ("crossing_clearance", F), ("capsule_clearance", F),
        ("prox_range", F), ("rab_range", F), ("rab_loss_probability", F), ("unit_scale", F),
        ("light_threshold", F), ("light_intensity", F), ("alpha", F),
        ("light_x", F), ("light_y", F), ("critic_radius", F),
        ("spawn_cx", F), ("spawn_cy", F), ("spawn_sx", F), ("spawn_sy", F), ("spawn_circle_radius", F),
        ("prox_threshold", F),
        ("cos_a", F * 8), ("sin_a", F * 8), ("rab_cos", F * 4), ("rab_sin", F * 4),
        ("ztilde_lut", F * N),
        ("face_nx", F * 12), ("face_ny", F * 12), ("face_px", F * 12), ("face_py", F * 12),
        ("seg_ax", F * MAX_SEG), ("seg_ay", F * MAX_SEG), ("seg_bx", F * MAX_SEG), ("seg_by", F * MAX_SEG),
        ("seg_sx", F * MAX_SEG), ("seg_sy", F * MAX_SEG),
        ("iw_ax", F * MAX_INTERNAL), ("iw_ay", F * MAX_INTERNAL),
        ("iw_tx", F * MAX_INTERNAL), ("iw_ty", F * MAX_INTERNAL),
        ("iw_nx", F * MAX_INTERNAL), ("iw_ny", F * MAX_INTERNAL), ("iw_len_sq", F * MAX_INTERNAL),
        ("gate", F * 12), ("zone", F * 12),
        ("mc_face_nx", F * 12), ("mc_face_ny", F * 12), ("mc_face_px", F * 12), ("mc_face_py", F * 12),
        ("mc_spawn_safe", F), ("mc_spawn_theta_max", F), ("mc_mode", I),
    ]


class SwarmState(C.Structure):
    _fields_ = [
        ("pos", C.c_void_p), ("yaw", C.c_void_p), ("prev_ground", C.c_void_p),
        ("cached_left", C.c_void_p), ("cached_right", C.c_void_p), ("fsm", C.c_void_p),
        ("beh_cache", C.c_void_p), ("mission_flags", C.c_void_p),
        ("episode_length_buf", C.c_void_p), ("episode_group_reward", C.c_void_p),
        ("completed_group_reward", C.c_void_p), ("completed_terminal_critic_state", C.c_void_p),
        ("scratch", C.c_void_p),
    ]


class SwarmNoise(C.Structure):
    _fields_ = [
        ("rab_u", C.c_void_p), ("turn_dur", C.c_void_p), ("spawn_u", C.c_void_p), ("yaw_u", C.c_void_p),
        ("spawn_rounds", I), ("seed", C.c_uint64), ("step_counter", C.c_uint64), ("env_offset", C.c_int64),
        ("rab_u2", C.c_void_p), ("mc_spawn_u", C.c_void_p),
        ("any_reset_mode", I), ("any_reset_bits", C.c_uint32),
    ]


class SwarmOut(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("reward", C.c_void_p), ("time_out", C.c_void_p), ("critic", C.c_void_p)]


# E-puck IR sensor bearings, SENS:28-37, and RAB projection axes, SENS:40-41.
_EPUCK_ANGLE_DIVISORS = (10.5884, 3.5999, 2.0, 1.2, 0.8571, 0.6667, 0.5806, 0.5247)


def sensor_tables():
    """float32 sensor direction tables computed with the same float32 torch ops as SENS:75-79."""
    ang = torch.tensor([math.pi / d for d in _EPUCK_ANGLE_DIVISORS], dtype=torch.float32)
    cos_a = torch.cos(ang)
    sin_a = -torch.sin(ang)
    rab = torch.tensor([45.0, 135.0, 225.0, 315.0], dtype=torch.float32) * (math.pi / 180.0)
    return cos_a.numpy(), sin_a.numpy(), torch.cos(rab).numpy(), torch.sin(rab).numpy()


def arena_segments(cfg):
    """ENV:615-628: dodecagon wall segments (ax, ay, bx, by) in double precision."""
    R, n = cfg.arena_circumradius, cfg.arena_num_sides
    verts = []
    for i in range(n):
        a = 2 * math.pi * i / n + math.pi / n
        verts.append((R * math.cos(a), R * math.sin(a)))
    return [(*verts[i], *verts[(i + 1) % n]) for i in range(n)]


def north_inradius(cfg):
    return cfg.arena_circumradius * math.cos(math.pi / cfg.arena_num_sides)  # ENV:649-650


def gate_south_y(cfg):
    return (north_inradius(cfg) - cfg.corridor_length) - cfg.gate_length  # ENV:652-656


def shelter_bounds(cfg):
    cx, cy = cfg.shelter_center
    sx, sy = cfg.shelter_size
    return cx - sx / 2, cx + sx / 2, cy - sy / 2, cy + sy / 2  # SHL:24-27


def internal_segments(cfg):
    """Mission-specific internal wall segments (ENV:630-645, SHL:29-35; none for XOR/HOM/FOR)."""
    m = cfg._mission
    if m == "dgt":
        hw = cfg.corridor_width / 2.0
        gs = gate_south_y(cfg)
        wl = cfg.side_wall_length
        return [(-hw, gs, -hw, gs + wl), (hw, gs, hw, gs + wl)]
    if m == "shl":
        left, right, bottom, top = shelter_bounds(cfg)
        return [(left, bottom, left, top), (right, bottom, right, top), (left, top, right, top)]
    return []


def build_params(cfg) -> SwarmParams:
    """Derive the kernel constants from an env cfg (any object with the CFG attributes)."""
    m = cfg._mission
    p = SwarmParams()
    p.abi_version = ABI_VERSION
    p.mission = MISSION_ID[m]
    if cfg.num_agents != N:
        raise ValueError(f"the fused step is specialised for {N} robots per env, got {cfg.num_agents}")
    p.obs_dim = 24 if (cfg.variant in ("dandelion", "daisy") or cfg.full_policy_observations) else 4
    p.discrete_actions = int(bool(cfg.discrete_actions))
    p.decimation = int(cfg.decimation)
    if p.decimation < 1:
        raise ValueError("decimation must be >= 1")
    p.max_episode_length = math.ceil(cfg.episode_length_s / (cfg.sim.dt * cfg.decimation))
    p.solver_iterations = max(1, int(getattr(cfg, "collision_solver_iterations", 4)))  # ENV:876
    p.has_light = int(bool(getattr(cfg, "has_light", True)))
    p.spawn_max_attempts = int(getattr(cfg, "spawn_max_attempts", 100))
    p.dt = cfg.sim.dt
    p.wheelbase = cfg.wheelbase
    p.max_wheel_speed = cfg.max_wheel_speed
    r = cfg.robot_radius
    eps = float(getattr(cfg, "wall_contact_epsilon", 1e-4))
    p.robot_radius = r
    p.robot_radius_sq = r ** 2            # SENS:274
    p.two_radius = 2 * r                  # ENV:1083
    p.wall_r_eff = r + 0.5 * float(getattr(cfg, "arena_wall_thickness", 0.01)) + eps  # ENV:1050-1054
    p.crossing_clearance = r + 0.5 * float(getattr(cfg, "shelter_wall_thickness", 0.0)) + eps  # ENV:909-913
    wall_t = float(getattr(cfg, "shelter_wall_thickness", getattr(cfg, "internal_wall_thickness", 0.01)))
    p.capsule_clearance = r + 0.5 * wall_t + eps  # ENV:981-990
    p.prox_range = cfg.prox_range
    p.rab_range = cfg.rab_range
    p.rab_loss_probability = cfg.rab_loss_probability
    p.unit_scale = cfg.unity_unit_scale_m
    p.light_threshold = cfg.light_threshold
    p.light_intensity = cfg.light_intensity
    p.alpha = cfg.alpha_parameter
    p.light_x, p.light_y = cfg.light_position[0], cfg.light_position[1]
    p.critic_radius = cfg.critic_state_radius
    p.spawn_cx, p.spawn_cy = cfg.spawn_area_center
    p.spawn_sx, p.spawn_sy = cfg.spawn_area_size
    p.spawn_circle_radius = float(getattr(cfg, "spawn_circle_radius", 0.0))
    p.prox_threshold = 0.1  # BEH:116 (the env never overrides it, ENV:98-102)

    cos_a, sin_a, rab_cos, rab_sin = sensor_tables()
    for k in range(8):
        p.cos_a[k], p.sin_a[k] = float(cos_a[k]), float(sin_a[k])
    for k in range(4):
        p.rab_cos[k], p.rab_sin[k] = float(rab_cos[k]), float(rab_sin[k])
    zt = 1.0 - 2.0 / (1.0 + torch.exp(torch.arange(N, dtype=torch.float32)))  # SENS:425 on float32 counts
    for k in range(N):
        p.ztilde_lut[k] = float(zt[k])

    arena = arena_segments(cfg)
    if len(arena) != 12:
        raise ValueError("the fused step is specialised for the 12-sided arena")
    for i, (ax, ay, bx, by) in enumerate(arena):  # ENV:849-872
        mx, my = 0.5 * (ax + bx), 0.5 * (ay + by)
        norm = math.sqrt(mx * mx + my * my) + 1e-12
        p.face_nx[i], p.face_ny[i] = -mx / norm, -my / norm
        p.face_px[i], p.face_py[i] = mx, my

    internal = internal_segments(cfg)
    segs = arena + internal
    p.n_segments, p.n_internal = len(segs), len(internal)
    seg32 = np.asarray(segs, dtype=np.float32).reshape(-1, 4)  # SENS:205 torch.tensor(..., float32)
    for i in range(len(segs)):
        p.seg_ax[i], p.seg_ay[i], p.seg_bx[i], p.seg_by[i] = (float(v) for v in seg32[i])
        p.seg_sx[i] = float(seg32[i, 2] - seg32[i, 0])  # float32 subtraction, SENS:212-213
        p.seg_sy[i] = float(seg32[i, 3] - seg32[i, 1])
    for i, (ax, ay, bx, by) in enumerate(internal):  # ENV:916-938
        abx, aby = bx - ax, by - ay
        length_sq = abx * abx + aby * aby
        length = math.sqrt(length_sq)
        p.iw_ax[i], p.iw_ay[i] = ax, ay
        p.iw_tx[i], p.iw_ty[i] = abx, aby
        p.iw_nx[i], p.iw_ny[i] = -aby / length, abx / length
        p.iw_len_sq[i] = length_sq

    ni = north_inradius(cfg)
    if m in ("dgt", "xor"):  # XOR inherits ENV:658-705 (xor_aggregation_env.py:62-64 only drops the segments)
        p.gate_mode = GATE_DGT
        gs = gate_south_y(cfg)
        p.gate[0] = cfg.corridor_width / 2.0
        p.gate[1] = gs
        p.gate[2] = gs + cfg.side_wall_length
    elif m == "shl":
        p.gate_mode = GATE_SHL
        left, right, bottom, top = shelter_bounds(cfg)
        t = cfg.shelter_wall_thickness
        vals = (left, right, bottom, top, r + t / 2, bottom - r, top + r, left - r, right + r)
        for i, v in enumerate(vals):
            p.gate[i] = v
    else:
        p.gate_mode = GATE_NONE

    if m == "dgt":  # ENV:720-745
        corr_south = ni - cfg.corridor_length
        vals = (cfg.gate_width / 2.0, corr_south - cfg.gate_length, corr_south, cfg.corridor_width / 2.0, ni)
    elif m == "xor":
        (c0, c1) = cfg.target_centers
        vals = (c0[0], c0[1], c1[0], c1[1], cfg.target_radius ** 2)
    elif m == "hom":
        vals = (cfg.goal_center[0], cfg.goal_center[1], 0.0, 0.0, cfg.goal_radius ** 2)
    elif m == "for":
        (c0, c1) = cfg.food_centers
        vals = (c0[0], c0[1], c1[0], c1[1], cfg.food_radius ** 2, cfg.food_radius, cfg.nest_top_y)
    else:
        (c0, c1) = cfg.black_area_centers
        vals = (c0[0], c0[1], c1[0], c1[1], cfg.black_area_radius ** 2, 0.0, 0.0) + shelter_bounds(cfg)
    for i, v in enumerate(vals):
        p.zone[i] = v
    return p


MC_TASK_MISSION = {
    "SwarmACB-DirectionalGate-v0": "dgt", "SwarmACB-XOR-v0": "xor", "SwarmACB-Homing-v0": "hom",
    "SwarmACB-Foraging-v0": "for", "SwarmACB-Sheltering-v0": "shl", "SwarmACB-SCA-v0": "shl", "SwarmACB-SHL-v0": "shl",
}


def build_mc_params(task: str = "SwarmACB-DirectionalGate-v0") -> SwarmParams:
    """Constants of scripts/manual_control.py's StandaloneDGTEnv (MC:99-210) for ``swarm_mc_tick``.

    Differences from the DirectMARLEnv classes that are reproduced: light at (0,-1.4) (MC:143), nest edge
    -0.63 (MC:162), no gate push-out for XOR (MC:469-470), arena faces derived from angles and resolved
    sequentially with r = 0.035 (MC:531-553), episode_steps = round(T/dt) (MC:123), polar spawn (MC:250-258).
    """
    from .cfg import MISSION_CFGS
    mission = MC_TASK_MISSION.get(task, "dgt")  # MC:70-79: unknown tasks fall back to dgt
    cfg = MISSION_CFGS[mission]()
    cfg.update_variant("daisy")                 # 24-dim observation (compute_obs_robot0), module-id actions
    cfg.light_position = (0.0, -1.4, 0.0)
    if mission == "for":
        cfg.nest_top_y = -0.63
    p = build_params(cfg)
    p.mc_mode = 1
    p.max_episode_length = int(round(cfg.episode_length_s / 0.1))
    r, n = 0.035, 12
    R = math.sqrt(2 * 4.91 / (12 * math.sin(2 * math.pi / 12)))
    inradius = R * math.cos(math.pi / n)
    for i in range(n):  # MC:536-544
        a1 = 2 * math.pi * i / n + math.pi / n
        a2 = 2 * math.pi * ((i + 1) % n) / n + math.pi / n
        mid = (a1 + a2) / 2.0
        p.mc_face_nx[i], p.mc_face_ny[i] = -math.cos(mid), -math.sin(mid)
        p.mc_face_px[i], p.mc_face_py[i] = inradius * math.cos(mid), inradius * math.sin(mid)
    p.mc_spawn_safe = inradius - r * 2
    p.mc_spawn_theta_max = math.pi if mission == "hom" else 2 * math.pi
    if mission == "xor":
        p.gate_mode = GATE_NONE
    if mission == "shl":  # MC:322-329 derives the bounds from float32 tensors
        half = np.asarray(cfg.shelter_size, dtype=np.float32) / np.float32(2.0)
        ctr = np.asarray(cfg.shelter_center, dtype=np.float32)
        left, right = float(ctr[0] - half[0]), float(ctr[0] + half[0])
        bottom, top = float(ctr[1] - half[1]), float(ctr[1] + half[1])
        t = cfg.shelter_wall_thickness
        for i, v in enumerate((left, right, bottom, top, r + t / 2, bottom - r, top + r, left - r, right + r)):
            p.gate[i] = v
        for i, v in enumerate((left, right, bottom, top)):
            p.zone[7 + i] = v
    return p


# ── fsm word packing (include/swarm_abi.h) ───────────────────────────────────────────────

def _enc_dir(d):
    d = torch.as_tensor(d)
    return torch.where(d > 0, 1, torch.where(d < 0, 2, 0)).to(torch.int32)


def _dec_dir(c):
    return torch.where(c == 1, 1.0, torch.where(c == 2, -1.0, 0.0)).to(torch.float32)


def pack_fsm(explore_state, explore_steps, explore_dir, photo_avoiding, photo_steps, photo_dir,
             anti_avoiding, anti_steps, anti_dir) -> torch.Tensor:
    def group(flag, steps, d):
        steps = torch.as_tensor(steps).to(torch.int32)
        if int(steps.min()) < 0 or int(steps.max()) > 7:
            raise ValueError("turn step counters must be in [0, 7]")
        return (torch.as_tensor(flag).to(torch.int32) & 1) | (steps << 1) | (_enc_dir(d) << 4)

    return (group(explore_state, explore_steps, explore_dir)
            | (group(photo_avoiding, photo_steps, photo_dir) << 6)
            | (group(anti_avoiding, anti_steps, anti_dir) << 12)).to(torch.int32)


def unpack_fsm(word: torch.Tensor) -> dict:
    out = {}
    for name, shift, flag_name, is_bool in (("explore", 0, "state", False), ("photo", 6, "avoiding", True),
                                            ("antiphoto", 12, "avoiding", True)):
        g = (word >> shift) & 63
        flag = g & 1
        out[f"_{name}_{flag_name}"] = flag.bool() if is_bool else flag.long()
        out[f"_{name}_steps"] = ((g >> 1) & 7).long()
        out[f"_{name}_dir"] = _dec_dir((g >> 4) & 3)
    return out
