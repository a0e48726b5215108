"""ctypes front-end of the CPU oracle (oracle/swarm_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this module; the product package never does.  State lives in host numpy arrays with the
same layouts as include/swarm_abi.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from swarmacb_isaaclab_b200.params import N, SwarmNoise, SwarmOut, SwarmParams, SwarmState

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libswarm_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    deps = (os.path.join(_HERE, "swarm_oracle.c"), os.path.join(_HERE, "..", "include", "swarm_abi.h"),
            os.path.join(_HERE, "..", "include", "swarm_detmath.h"))
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(d) for d in deps):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        for name in ("swarm_oracle_step", "swarm_oracle_reset", "swarm_oracle_critic_state", "swarm_oracle_mc_tick",
                     "swarm_oracle_mc_reset", "swarm_oracle_set_threads"):
            getattr(_lib, name).restype = C.c_int
    return _lib


STATE_SPEC = {  # name -> (per-env shape, dtype)
    "pos": ((N, 2), np.float32), "yaw": ((N,), np.float32), "prev_ground": ((N,), np.float32),
    "cached_left": ((N,), np.float32), "cached_right": ((N,), np.float32), "fsm": ((N,), np.int32),
    "beh_cache": ((6, N), np.float32), "mission_flags": ((N,), np.uint8),
    "episode_length_buf": ((), np.int64), "episode_group_reward": ((), np.float32),
    "completed_group_reward": ((), np.float32), "completed_terminal_critic_state": ((N, 5), np.float32),
}


def new_state(E: int) -> dict:
    s = {k: np.zeros((E,) + shp, dtype=dt) for k, (shp, dt) in STATE_SPEC.items()}
    s["prev_ground"][:] = 0.5
    return s


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _state_struct(state: dict) -> SwarmState:
    st = SwarmState()
    for k in STATE_SPEC:
        a = state[k]
        assert a.flags.c_contiguous and a.dtype == STATE_SPEC[k][1], k
        setattr(st, k, _ptr(a))
    return st


def _noise_struct(E, rab_u=None, turn_dur=None, spawn_u=None, yaw_u=None):
    nz = SwarmNoise()
    keep = []
    if rab_u is not None:
        rab_u = np.ascontiguousarray(rab_u, dtype=np.float32).reshape(E, N, N)
        nz.rab_u = _ptr(rab_u)
    if turn_dur is not None:
        turn_dur = np.ascontiguousarray(turn_dur, dtype=np.int32).reshape(E, N, 3)
        nz.turn_dur = _ptr(turn_dur)
    if spawn_u is not None:
        spawn_u = np.ascontiguousarray(spawn_u, dtype=np.float32).reshape(-1, E, N, 2)
        nz.spawn_u = _ptr(spawn_u)
        nz.spawn_rounds = spawn_u.shape[0]
    if yaw_u is not None:
        yaw_u = np.ascontiguousarray(yaw_u, dtype=np.float32).reshape(E, N)
        nz.yaw_u = _ptr(yaw_u)
    keep.extend([rab_u, turn_dur, spawn_u, yaw_u])
    return nz, keep


def step(params: SwarmParams, state: dict, actions: np.ndarray, *, rab_u, turn_dur=None, spawn_u=None,
         yaw_u=None):
    """One env.step on the host.  Mutates ``state``; returns (obs, reward, time_out)."""
    E = state["pos"].shape[0]
    if params.discrete_actions:
        actions = np.ascontiguousarray(actions, dtype=np.int64).reshape(E, N)
        if turn_dur is None:
            raise ValueError("discrete actions need injected turn_dur")
    else:
        actions = np.ascontiguousarray(actions, dtype=np.float32).reshape(E, N, 2)
    if spawn_u is None:
        spawn_u = np.zeros((1, E, N, 2), np.float32)
    if yaw_u is None:
        yaw_u = np.zeros((E, N), np.float32)
    obs = np.zeros((E, N, params.obs_dim), np.float32)
    reward = np.zeros((E,), np.float32)
    time_out = np.zeros((E,), np.uint8)
    out = SwarmOut(_ptr(obs), _ptr(reward), _ptr(time_out))
    st = _state_struct(state)
    nz, _keep = _noise_struct(E, rab_u, turn_dur, spawn_u, yaw_u)
    rc = lib().swarm_oracle_step(C.byref(params), C.byref(st), _ptr(actions), C.byref(nz), C.byref(out), E)
    if rc != 0:
        raise RuntimeError(f"swarm_oracle_step failed: {rc}")
    return obs, reward, time_out.astype(bool)


def reset(params: SwarmParams, state: dict, *, rab_u, spawn_u, yaw_u):
    E = state["pos"].shape[0]
    obs = np.zeros((E, N, params.obs_dim), np.float32)
    out = SwarmOut(_ptr(obs), None, None)
    st = _state_struct(state)
    nz, _keep = _noise_struct(E, rab_u, None, spawn_u, yaw_u)
    rc = lib().swarm_oracle_reset(C.byref(params), C.byref(st), C.byref(nz), C.byref(out), E)
    if rc != 0:
        raise RuntimeError(f"swarm_oracle_reset failed: {rc}")
    return obs


def critic_state(params: SwarmParams, state: dict) -> np.ndarray:
    E = state["pos"].shape[0]
    out = np.zeros((E, N, 5), np.float32)
    st = _state_struct(state)
    rc = lib().swarm_oracle_critic_state(C.byref(params), C.byref(st), _ptr(out), E)
    if rc != 0:
        raise RuntimeError(f"swarm_oracle_critic_state failed: {rc}")
    return out


def detmath_sincos(a: np.ndarray):
    a = np.ascontiguousarray(a, dtype=np.float32)
    sn, cs = np.empty_like(a), np.empty_like(a)
    lib().swarm_oracle_sincos(a.size, _ptr(a), _ptr(sn), _ptr(cs))
    return sn, cs


def detmath_atan2(y: np.ndarray, x: np.ndarray):
    y = np.ascontiguousarray(y, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(y)
    lib().swarm_oracle_atan2(y.size, _ptr(y), _ptr(x), _ptr(out))
    return out


def set_threads(n: int | None = None) -> int:
    """Use ``n`` OpenMP threads (default: every core this process may run on); returns the count in effect."""
    if n is None:
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return int(lib().swarm_oracle_set_threads(int(n)))


MC_PRE, MC_PHYSICS, MC_POST = 1, 2, 4


def mc_tick(params: SwarmParams, state: dict, module_ids, wheels, *, rab_u=None, rab_u2=None, turn_dur=None,
            mc_spawn_u=None, flags=MC_PRE | MC_PHYSICS | MC_POST):
    """One manual-control tick (scripts/manual_control.py:721-757) on the host; returns (obs24, reward, rolled_over)."""
    E = state["pos"].shape[0]
    ids = np.ascontiguousarray(module_ids, dtype=np.int64).reshape(E, N) if module_ids is not None else None
    wh = np.ascontiguousarray(wheels, dtype=np.float32).reshape(E, N, 2) if wheels is not None else None
    obs = np.zeros((E, N, 24), np.float32)
    reward = np.zeros((E,), np.float32)
    time_out = np.zeros((E,), np.uint8)
    out = SwarmOut(_ptr(obs), _ptr(reward), _ptr(time_out))
    st = _state_struct(state)
    nz, _keep = _noise_struct(E, rab_u, turn_dur, None, None)
    extra = []
    if rab_u2 is not None:
        a = np.ascontiguousarray(rab_u2, dtype=np.float32).reshape(E, N, N)
        nz.rab_u2 = _ptr(a)
        extra.append(a)
    if mc_spawn_u is None:
        mc_spawn_u = np.zeros((E, N, 3), np.float32)
    b = np.ascontiguousarray(mc_spawn_u, dtype=np.float32).reshape(E, N, 3)
    nz.mc_spawn_u = _ptr(b)
    rc = lib().swarm_oracle_mc_tick(C.byref(params), C.byref(st), _ptr(ids), _ptr(wh), C.byref(nz), C.byref(out),
                                    int(flags), E)
    if rc != 0:
        raise RuntimeError(f"swarm_oracle_mc_tick failed: {rc}")
    return obs, reward, time_out.astype(bool)


def mc_reset(params: SwarmParams, state: dict, mc_spawn_u):
    E = state["pos"].shape[0]
    st = _state_struct(state)
    nz = SwarmNoise()
    b = np.ascontiguousarray(mc_spawn_u, dtype=np.float32).reshape(E, N, 3)
    nz.mc_spawn_u = _ptr(b)
    rc = lib().swarm_oracle_mc_reset(C.byref(params), C.byref(st), C.byref(nz), E)
    if rc != 0:
        raise RuntimeError(f"swarm_oracle_mc_reset failed: {rc}")
