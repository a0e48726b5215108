"""Viewer feed from device state.

The reference draws one environment with Omniverse markers (`directional_gate_env.py:393-609`) or with its
stand-alone viewers (`scripts/manual_control.py:692-980`, `scripts/manual_control_isaac.py:228-809`): arena,
internal walls, mission zones, light, robots with their heading, the IR rays and the range-and-bearing disc of one
selected robot.  Everything those viewers need is either static per mission (`scene()`) or a 20-robot slice of the
device state plus the last observation (`frame()`): a frame costs a handful of sub-kilobyte device-to-host copies
of ONE environment, never the batch.  `render_svg()` turns both into a picture without any GUI dependency
(pygame / Omniverse are not part of this image); a GUI would consume the same two dicts.
"""
from __future__ import annotations

import math

import numpy as np

from .params import FSM_STATE_MASK, N, unpack_fsm

MODULE_NAMES = ("stop", "exploration", "attraction", "repulsion", "phototaxis", "anti-phototaxis")   # BEH:34-41


def scene(env) -> dict:
    """Static geometry of the env's mission, from the same SwarmParams block the kernel reads."""
    p = env.params
    ns = int(p.n_segments)
    seg = np.array([[p.seg_ax[i], p.seg_ay[i], p.seg_bx[i], p.seg_by[i]] for i in range(ns)], np.float32)
    z = [float(v) for v in p.zone]
    zones = []
    if p.mission == 0:     # ENV:707-750: white gate strip, black corridor
        zones += [{"kind": "rect", "colour": "white", "x0": -z[0], "x1": z[0], "y0": z[1], "y1": z[2]},
                  {"kind": "rect", "colour": "black", "x0": -z[3], "x1": z[3], "y0": z[2], "y1": z[4]}]
    elif p.mission == 1:   # XOR:110-124 two black targets
        zones += [{"kind": "circle", "colour": "black", "cx": z[0], "cy": z[1], "r": math.sqrt(z[4])},
                  {"kind": "circle", "colour": "black", "cx": z[2], "cy": z[3], "r": math.sqrt(z[4])}]
    elif p.mission == 2:   # HOM:76-85 black goal
        zones += [{"kind": "circle", "colour": "black", "cx": z[0], "cy": z[1], "r": math.sqrt(z[4])}]
    elif p.mission == 3:   # FOR:104-125 black food discs, white nest below nest_top_y
        zones += [{"kind": "circle", "colour": "black", "cx": z[0], "cy": z[1], "r": math.sqrt(z[4])},
                  {"kind": "circle", "colour": "black", "cx": z[2], "cy": z[3], "r": math.sqrt(z[4])},
                  {"kind": "halfplane_below", "colour": "white", "y": z[6]}]
    else:                  # SHL:106-122 black discs, white shelter
        zones += [{"kind": "circle", "colour": "black", "cx": z[0], "cy": z[1], "r": math.sqrt(z[4])},
                  {"kind": "circle", "colour": "black", "cx": z[2], "cy": z[3], "r": math.sqrt(z[4])},
                  {"kind": "rect", "colour": "white", "x0": z[7], "x1": z[8], "y0": z[9], "y1": z[10]}]
    return {
        "mission": int(p.mission), "arena_faces": seg[:12], "internal_walls": seg[12:],
        "zones": zones, "light": (float(p.light_x), float(p.light_y)) if p.has_light else None,
        "robot_radius": float(p.robot_radius), "prox_range": float(p.prox_range), "rab_range": float(p.rab_range),
        "sensor_cos": np.array(list(p.cos_a), np.float32), "sensor_sin": np.array(list(p.sin_a), np.float32),
        "obs_dim": int(p.obs_dim), "discrete_actions": bool(p.discrete_actions),
    }


def frame(env, env_index: int = 0) -> dict:
    """Dynamic state of ONE environment as small host arrays (device -> host copies of that env's rows only)."""
    e = int(env_index)
    if not 0 <= e < env.num_envs:
        raise IndexError(f"env_index {e} out of range for {env.num_envs} envs")
    fsm = env._fsm[e].detach().cpu()
    out = {
        "env_index": e,
        "pos": env.agent_pos[e].detach().cpu().numpy().copy(),                 # (N,2)
        "yaw": env.agent_yaw[e].detach().cpu().numpy().copy(),                 # (N,)
        "ground": env.prev_ground_color[e].detach().cpu().numpy().copy(),      # (N,) 0 black / .5 grey / 1 white
        "wheels": np.stack([env._cached_left_vel[e].detach().cpu().numpy(),
                            env._cached_right_vel[e].detach().cpu().numpy()], -1),   # (N,2) m/s
        "behaviour": {k: v.numpy() for k, v in unpack_fsm(fsm & FSM_STATE_MASK).items()},
        "obs": env._obs[e].detach().cpu().numpy().copy(),                      # (N,obs_dim), the last observation
        "episode_step": int(env.episode_length_buf[e]),
        "episode_group_reward": float(env._episode_group_reward[e]),
        "last_reward": float(env._reward[e]),
    }
    if int(env.params.mission) == 3:
        out["has_food"] = (env._mission_flags[e].detach().cpu().numpy() & 1).astype(bool)
    return out


def render_svg(sc: dict, fr: dict, size: int = 640, selected: int = 0) -> str:
    """One environment as an SVG string: zones, walls, light, robots with headings, and the IR rays / RAB disc of the
    `selected` robot (what ENV:393-609's sensor markers show).  Pure string formatting."""
    R = float(np.abs(sc["arena_faces"][:, :2]).max()) * 1.08
    s = size / (2 * R)

    def X(x):
        return (x + R) * s

    def Y(y):
        return (R - y) * s

    grey = {"black": "#222", "white": "#fafafa"}
    parts = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{size}" height="{size}" viewBox="0 0 {size} {size}">',
             f'<rect width="{size}" height="{size}" fill="#777"/>']
    poly = " ".join(f"{X(a):.1f},{Y(b):.1f}" for a, b in sc["arena_faces"][:, :2])
    parts.append(f'<polygon points="{poly}" fill="#999" stroke="#333" stroke-width="3"/>')
    for z in sc["zones"]:
        if z["kind"] == "circle":
            parts.append(f'<circle cx="{X(z["cx"]):.1f}" cy="{Y(z["cy"]):.1f}" r="{z["r"] * s:.1f}" fill="{grey[z["colour"]]}"/>')
        elif z["kind"] == "rect":
            parts.append(f'<rect x="{X(z["x0"]):.1f}" y="{Y(z["y1"]):.1f}" width="{(z["x1"] - z["x0"]) * s:.1f}" '
                         f'height="{(z["y1"] - z["y0"]) * s:.1f}" fill="{grey[z["colour"]]}"/>')
        else:
            parts.append(f'<rect x="0" y="{Y(z["y"]):.1f}" width="{size}" height="{size - Y(z["y"]):.1f}" '
                         f'fill="{grey[z["colour"]]}" opacity="0.55"/>')
    for ax, ay, bx, by in sc["internal_walls"]:
        parts.append(f'<line x1="{X(ax):.1f}" y1="{Y(ay):.1f}" x2="{X(bx):.1f}" y2="{Y(by):.1f}" stroke="#40210f" stroke-width="4"/>')
    if sc["light"] is not None:
        lx, ly = sc["light"]
        parts.append(f'<circle cx="{X(lx):.1f}" cy="{min(max(Y(ly), 6), size - 6):.1f}" r="6" fill="#ffd400" stroke="#a80"/>')
    r = sc["robot_radius"]
    sel = int(selected) % N
    px, py, th = float(fr["pos"][sel, 0]), float(fr["pos"][sel, 1]), float(fr["yaw"][sel])
    parts.append(f'<circle cx="{X(px):.1f}" cy="{Y(py):.1f}" r="{sc["rab_range"] * s:.1f}" fill="none" stroke="#3a6" '
                 f'stroke-dasharray="4 4"/>')
    if sc["obs_dim"] == 24:   # the selected robot's 8 IR rays, shortened by its proximity readings (SENS:85-142)
        for k in range(8):
            ca, sa = float(sc["sensor_cos"][k]), float(sc["sensor_sin"][k])
            dx, dy = ca * math.cos(th) - sa * math.sin(th), ca * math.sin(th) + sa * math.cos(th)
            reach = sc["prox_range"] * (1.0 - float(fr["obs"][sel, k]))
            colour = "#e33" if fr["obs"][sel, k] > 0 else "#8cf"
            parts.append(f'<line x1="{X(px):.1f}" y1="{Y(py):.1f}" x2="{X(px + dx * reach):.1f}" y2="{Y(py + dy * reach):.1f}" '
                         f'stroke="{colour}" stroke-width="1.5"/>')
    for i in range(N):
        x, y, a = float(fr["pos"][i, 0]), float(fr["pos"][i, 1]), float(fr["yaw"][i])
        fill = "#1e6fd9" if i != sel else "#e08a00"
        if "has_food" in fr and bool(fr["has_food"][i]):
            fill = "#2a9d3a"
        parts.append(f'<circle class="robot" cx="{X(x):.1f}" cy="{Y(y):.1f}" r="{r * s:.1f}" fill="{fill}" stroke="#000"/>')
        parts.append(f'<line x1="{X(x):.1f}" y1="{Y(y):.1f}" x2="{X(x + r * math.cos(a)):.1f}" y2="{Y(y + r * math.sin(a)):.1f}" '
                     f'stroke="#fff" stroke-width="2"/>')
    parts.append(f'<text x="8" y="18" font-family="monospace" font-size="13" fill="#fff">env {fr["env_index"]}  '
                 f'step {fr["episode_step"]}  return {fr["episode_group_reward"]:.0f}</text>')
    parts.append("</svg>")
    return "\n".join(parts)
