"""``StandaloneSwarmEnv`` - the reference's no-Isaac kinematic path (scripts/manual_control.py ``StandaloneDGTEnv``,
BASELINE config 1) on the fused CUDA kernel.

The reference's pygame loop does, per 10 Hz tick (manual_control.py:721-757): sensors at the current pose ->
``BehaviorModules.dispatch`` for robots 1..19 (robot 0 is driven by the keyboard) -> ``step(left, right)`` ->
episode roll-over -> ``compute_obs_robot0()`` (a second range-and-bearing noise draw).  ``tick()`` runs that whole
sequence in one kernel launch; ``step()`` and ``compute_obs_robot0()`` expose the reference's two methods on their
own.  Its step is NOT the DirectMARLEnv step: single Gauss-Seidel wall pass with r = 0.035, one gate pass, one robot
pass, no iterative solver, polar spawn, light at (0, -1.4) (last row).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .params import FSM_STATE_MASK, N, SwarmNoise, SwarmOut, SwarmState, build_mc_params, unpack_fsm

MC_PRE, MC_PHYSICS, MC_POST = 1, 2, 4


class StandaloneSwarmEnv:
    def __init__(self, num_agents: int = N, device: str = "cuda:0", task: str = "SwarmACB-DirectionalGate-v0",
                 num_envs: int = 1, seed: int = 0, env_offset: int = 0):
        if num_agents != N:
            raise ValueError(f"the fused step is specialised for {N} robots")
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("StandaloneSwarmEnv runs only on a CUDA device (no CPU fallback)")
        self._lib = _lib.load()
        self.task = task
        self.params = build_mc_params(task)
        self.E, self.N = int(num_envs), N
        self.episode_steps = int(self.params.max_episode_length)
        E, dev = self.E, self.device
        f32 = dict(dtype=torch.float32, device=dev)
        self.pos = torch.zeros(E, N, 2, **f32)
        self.yaw = torch.zeros(E, N, **f32)
        self.prev_ground_color = torch.full((E, N), 0.5, **f32)
        self._fsm = torch.zeros(E, N, dtype=torch.int32, device=dev)
        self._mission_flags = torch.zeros(E, N, dtype=torch.uint8, device=dev)
        self.step_count = torch.zeros(E, dtype=torch.long, device=dev)
        self.episode_reward = torch.zeros(E, **f32)
        self.completed_episode_reward = torch.zeros(E, **f32)
        self.step_reward = torch.zeros(E, **f32)
        self._rolled = torch.zeros(E, dtype=torch.uint8, device=dev)
        self._obs = torch.zeros(E, N, 24, **f32)
        self._zeros = torch.zeros(E, N, **f32)
        self._state = SwarmState(
            self.pos.data_ptr(), self.yaw.data_ptr(), self.prev_ground_color.data_ptr(), self._zeros.data_ptr(),
            self._zeros.data_ptr(), self._fsm.data_ptr(), None, self._mission_flags.data_ptr(), self.step_count.data_ptr(),
            self.episode_reward.data_ptr(), self.completed_episode_reward.data_ptr(), None, None)
        self._out = SwarmOut(self._obs.data_ptr(), self.step_reward.data_ptr(), self._rolled.data_ptr())
        self._seed, self._counter, self._env_offset = int(seed), 0, int(env_offset)
        self._injected: dict = {}
        self.reset()

    @property
    def behavior_state(self) -> dict:
        return unpack_fsm(self._fsm)

    @property
    def has_food(self):
        return (self._mission_flags & 1).bool()

    @property
    def prev_in_nest(self):
        return ((self._mission_flags >> 1) & 1).bool()

    def inject_noise(self, rab_u=None, rab_u2=None, turn_dur=None, mc_spawn_u=None):
        """Parity mode: draws for the NEXT call (packet-loss uniforms of both sensor passes, turn durations, spawn)."""
        E, dev = self.E, self.device
        shapes = {"rab_u": (E, N, N), "rab_u2": (E, N, N), "turn_dur": (E, N, 3), "mc_spawn_u": (E, N, 3)}
        inj = {}
        for k, v in (("rab_u", rab_u), ("rab_u2", rab_u2), ("turn_dur", turn_dur), ("mc_spawn_u", mc_spawn_u)):
            if v is not None:
                dt = torch.int32 if k == "turn_dur" else torch.float32
                inj[k] = torch.as_tensor(v, device=dev).to(dt).reshape(shapes[k]).contiguous()
        self._injected = inj

    def _noise(self) -> SwarmNoise:
        inj, self._injected = self._injected, {}
        self._keep = inj
        nz = SwarmNoise()
        for k, t in inj.items():
            setattr(nz, k, t.data_ptr())
        nz.seed, nz.step_counter, nz.env_offset = self._seed, self._counter, self._env_offset
        self._counter += 1
        return nz

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self):
        """MC:245-269 (polar spawn; episode statistics cleared)."""
        nz = self._noise()
        with torch.cuda.device(self.device):
            rc = self._lib.swarm_mc_reset(C.byref(self.params), C.byref(self._state), C.byref(nz), self.E, self._stream())
        _lib.check(rc, "swarm_mc_reset")

    def _call(self, flags, module_ids=None, wheels=None):
        ids = None if module_ids is None else torch.as_tensor(module_ids, device=self.device).to(torch.long).reshape(self.E, N).contiguous()
        wh = None if wheels is None else torch.as_tensor(wheels, device=self.device).to(torch.float32).reshape(self.E, N, 2).contiguous()
        nz = self._noise()
        with torch.cuda.device(self.device):
            rc = self._lib.swarm_mc_tick(C.byref(self.params), C.byref(self._state),
                                         C.c_void_p(ids.data_ptr()) if ids is not None else None,
                                         C.c_void_p(wh.data_ptr()) if wh is not None else None,
                                         C.byref(nz), C.byref(self._out), int(flags), self.E, self._stream())
        _lib.check(rc, "swarm_mc_tick")

    def tick(self, module_ids, robot0_left=0.0, robot0_right=0.0):
        """One full loop iteration (MC:721-757).  module_ids (E,N) int; robot 0's wheel command in m/s.
        Returns (obs24 (E,N,24), step_reward (E), rolled_over (E) bool)."""
        wheels = torch.zeros(self.E, N, 2, dtype=torch.float32, device=self.device)
        wheels[:, 0, 0] = torch.as_tensor(robot0_left, dtype=torch.float32, device=self.device)
        wheels[:, 0, 1] = torch.as_tensor(robot0_right, dtype=torch.float32, device=self.device)
        self._call(MC_PRE | MC_PHYSICS | MC_POST, module_ids, wheels)
        return self._obs, self.step_reward, self._rolled.view(torch.bool)

    def step(self, left_vel, right_vel):
        """MC:355-423 ``StandaloneDGTEnv.step(left_vel, right_vel)`` with (E,N) wheel speeds in m/s (+ roll-over)."""
        wheels = torch.stack([torch.as_tensor(left_vel, device=self.device).reshape(self.E, N),
                              torch.as_tensor(right_vel, device=self.device).reshape(self.E, N)], dim=-1)
        self._call(MC_PHYSICS, None, wheels)

    def compute_obs_robot0(self) -> dict:
        """MC:425-465: sensor breakdown of robot 0 (draws a fresh packet-loss sample)."""
        self._call(MC_POST)
        o = self._obs[0, 0].tolist()
        return {"prox_8": o[0:8], "light_8": o[8:16], "ground_3": o[16:19], "ztilde": o[19], "rab_4": o[20:24], "obs_24": o}

    def load_state(self, state: dict):
        m = {"pos": self.pos, "yaw": self.yaw, "prev_ground": self.prev_ground_color, "fsm": self._fsm,
             "mission_flags": self._mission_flags, "episode_length_buf": self.step_count,
             "episode_group_reward": self.episode_reward}
        for k, dst in m.items():
            dst.copy_(torch.as_tensor(state[k]).to(dst.dtype).reshape(dst.shape))

    def dump_state(self) -> dict:
        m = {"pos": self.pos, "yaw": self.yaw, "prev_ground": self.prev_ground_color, "fsm": self._fsm,
             "mission_flags": self._mission_flags, "episode_length_buf": self.step_count,
             "episode_group_reward": self.episode_reward}
        out = {k: v.detach().cpu().numpy().copy() for k, v in m.items()}
        out["fsm"] &= FSM_STATE_MASK   # bits 18..23 hold pre-drawn turn-duration bits, not reference state
        return out
