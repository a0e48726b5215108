"""``SwarmEnv`` - the reference's DirectMARLEnv-style task API on top of the fused CUDA step.

Host-side mirror of the duck-typed protocol the reference trainers and scripts rely on
(agents/poca_trainer.py:204-222,376,509,575,619 of the reference):

* ``reset() -> (obs_dict, info)``, ``step(action_dict) -> (obs, reward, terminated, truncated, info)``
  with per-agent dicts keyed ``epuck_0..epuck_19``; ``obs_dict[a]`` is an ``(E, obs_dim)`` view.
* ``unwrapped`` -> self, ``cfg``, ``device``, ``num_envs``, ``scene.num_envs``, ``max_episode_length``,
  ``episode_length_buf``, ``get_critic_state()``, ``completed_terminal_critic_state``,
  ``completed_group_reward``, ``agent_pos``, ``agent_yaw``, ``prev_ground_color``.
* the seven Gymnasium ids of missions/*/__init__.py via :func:`make` (and ``gym.register`` when
  gymnasium is importable), each with the ``env_cfg_entry_point`` kwarg.

All state lives in torch-owned device tensors; the extension borrows raw pointers for the duration
of a call and enqueues on torch's current stream without synchronising.  There is no CPU fallback:
constructing the env without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import math
import types

import torch

from . import _lib
from .cfg import TASK_CFGS, DirectionalGateEnvCfg
from .params import FSM_STATE_MASK, N, SwarmNoise, SwarmOut, SwarmState, build_params, pack_fsm, unpack_fsm

_AGENTS = [f"epuck_{i}" for i in range(N)]


class SwarmEnv:
    """Batched 20-e-puck swarm environment (five missions, five CASA variants)."""

    metadata = {"render_modes": [None]}

    CRITIC_POOL = 4   # buffers get_critic_state() hands out in turn when the critic state is fused into the step

    def __init__(self, cfg: DirectionalGateEnvCfg | None = None, render_mode: str | None = None,
                 env_offset: int = 0, fused_critic: bool = True, **kwargs):
        if cfg is None:
            cfg = DirectionalGateEnvCfg()
        self.cfg = cfg
        self.render_mode = render_mode
        self.device = self._resolve_device(cfg.sim.device)
        self._lib = self._load_library()
        self.params = build_params(cfg)
        self.num_envs = int(cfg.scene.num_envs)
        self.scene = types.SimpleNamespace(num_envs=self.num_envs)
        self.sim = types.SimpleNamespace(has_gui=lambda: False, device=str(self.device))
        self.max_episode_length = math.ceil(cfg.episode_length_s / (cfg.sim.dt * cfg.decimation))
        self.possible_agents = list(cfg.possible_agents)
        self.agents = list(cfg.possible_agents)
        self.num_agents = N
        self.extras: dict = {}

        E, dev = self.num_envs, self.device
        f32 = dict(dtype=torch.float32, device=dev)
        self._agent_pos = torch.zeros(E, N, 2, **f32)
        self._agent_yaw = torch.zeros(E, N, **f32)
        self.prev_ground_color = torch.full((E, N), 0.5, **f32)
        self._cached_left_vel = torch.zeros(E, N, **f32)
        self._cached_right_vel = torch.zeros(E, N, **f32)
        self._fsm = torch.zeros(E, N, dtype=torch.int32, device=dev)
        self._beh_cache = torch.zeros(E, 6, N, **f32)
        self._mission_flags = torch.zeros(E, N, dtype=torch.uint8, device=dev)
        self._episode_length_buf = torch.zeros(E, dtype=torch.long, device=dev)
        self._episode_group_reward = torch.zeros(E, **f32)
        self.completed_group_reward = torch.zeros(E, **f32)
        self.completed_terminal_critic_state = torch.zeros(E, N, 5, **f32)
        self._scratch = torch.zeros(8, dtype=torch.int32, device=dev)
        self.reset_buf = torch.zeros(E, dtype=torch.bool, device=dev)

        self.obs_dim = int(self.params.obs_dim)
        self.act_dim = 1 if self.params.discrete_actions else 2
        self._obs = torch.zeros(E, N, self.obs_dim, **f32)
        self._reward = torch.zeros(E, **f32)
        self._time_out = torch.zeros(E, dtype=torch.uint8, device=dev)
        # get_critic_state() fused into the step: see get_critic_state()
        self.fused_critic = bool(fused_critic)
        self._critic_pool = [torch.zeros(E, N, 5, **f32) for _ in range(self.CRITIC_POOL)] if fused_critic else []
        self._critic_slot = 0          # pool buffer the next fused write goes to
        self._critic_fresh = False     # that buffer holds get_critic_state() of the CURRENT state
        self._critic_last = None       # tensor handed out for the CURRENT state (asked twice -> same tensor, no launch)
        self._critic_period = 0        # learned number of env.steps between two get_critic_state() calls
        self._steps_since_critic = 0
        self._terminated = torch.zeros(E, dtype=torch.bool, device=dev)
        self._act_buf = torch.zeros(E, N, self.act_dim, dtype=torch.long if self.params.discrete_actions else torch.float32,
                                    device=dev)

        self._state = SwarmState(
            self.agent_pos.data_ptr(), self.agent_yaw.data_ptr(), self.prev_ground_color.data_ptr(),
            self._cached_left_vel.data_ptr(), self._cached_right_vel.data_ptr(), self._fsm.data_ptr(),
            self._beh_cache.data_ptr(), self._mission_flags.data_ptr(), self.episode_length_buf.data_ptr(),
            self._episode_group_reward.data_ptr(), self.completed_group_reward.data_ptr(),
            self.completed_terminal_critic_state.data_ptr(), self._scratch.data_ptr())
        self._out = SwarmOut(self._obs.data_ptr(), self._reward.data_ptr(), self._time_out.data_ptr(), None)
        self._out_c = SwarmOut(self._obs.data_ptr(), self._reward.data_ptr(), self._time_out.data_ptr(), None)
        self._seed = int(cfg.seed) if getattr(cfg, "seed", None) is not None else 0
        self._step_counter = 0
        self._env_offset = int(env_offset)
        self._injected: dict = {}
        self.job_reset_clock = None    # sharding.JobResetClock: job-wide ENV:1262 coupling for de-synchronised shards
        self._obs_views = {a: self._obs[:, i] for i, a in enumerate(self.possible_agents)}
        self._len_version = self.episode_length_buf._version
        self._pose_version = (self._agent_pos._version, self._agent_yaw._version)

    # The kernels hold raw pointers to these tensors: rebinding the attribute (``env.episode_length_buf = t``) must
    # not detach the caller's tensor from them, so assignment copies INTO the tensor the kernel sees.
    @property
    def episode_length_buf(self) -> torch.Tensor:
        return self._episode_length_buf

    @episode_length_buf.setter
    def episode_length_buf(self, value):
        self._episode_length_buf.copy_(torch.as_tensor(value, device=self.device))

    @property
    def agent_pos(self) -> torch.Tensor:
        return self._agent_pos

    @agent_pos.setter
    def agent_pos(self, value):
        self._agent_pos.copy_(torch.as_tensor(value, device=self.device))

    @property
    def agent_yaw(self) -> torch.Tensor:
        return self._agent_yaw

    @agent_yaw.setter
    def agent_yaw(self, value):
        self._agent_yaw.copy_(torch.as_tensor(value, device=self.device))

    # ── device / library binding ────────────────────────────────────────────────────────────
    def _resolve_device(self, name) -> torch.device:
        """The step exists only as the CUDA kernel: anything but a usable CUDA device is an error."""
        device = torch.device(name)
        if device.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError(
                "SwarmEnv runs only on a CUDA device (the fused sm_100a step has no CPU fallback); "
                f"cfg.sim.device={name!r}, torch.cuda.is_available()={torch.cuda.is_available()}")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        return device

    def _load_library(self):
        return _lib.load()  # raises SwarmLibraryError when libswarmstep.so is missing

    def _device_guard(self):
        return torch.cuda.device(self.device)

    # ── protocol ────────────────────────────────────────────────────────────────────────────
    @property
    def unwrapped(self):
        return self

    def _spaces(self):
        """Per-agent spaces the way isaaclab's DirectMARLEnv derives them from the cfg's integer sizes (gymnasium
        objects when gymnasium is installed, the cfg's integers otherwise)."""
        obs, act = dict(self.cfg.observation_spaces), dict(self.cfg.action_spaces)
        try:
            import gymnasium as gym
            import numpy as np
        except ImportError:
            return obs, act
        obs = {a: gym.spaces.Box(-np.inf, np.inf, (int(n),), np.float32) for a, n in obs.items()}
        if self.params.discrete_actions:
            act = {a: gym.spaces.Discrete(int(self.cfg.num_actions)) for a in act}
        else:
            act = {a: gym.spaces.Box(-1.0, 1.0, (int(n),), np.float32) for a, n in act.items()}
        return obs, act

    @property
    def observation_spaces(self):
        return self._spaces()[0]

    @property
    def action_spaces(self):
        return self._spaces()[1]

    def observation_space(self, agent):
        return self.observation_spaces[agent]

    def action_space(self, agent):
        return self.action_spaces[agent]

    @property
    def _has_food(self):  # FOR:36
        return (self._mission_flags & 1).bool()

    @property
    def _prev_in_nest(self):  # FOR:37
        return ((self._mission_flags >> 1) & 1).bool()

    @property
    def behavior_state(self) -> dict:
        """Unpacked behaviour-module state machines (BEH:141-153 field names)."""
        return unpack_fsm(self._fsm)

    def set_behavior_state(self, **fields):
        cur = self.behavior_state
        cur.update({k: torch.as_tensor(v, device=self.device) for k, v in fields.items()})
        self._fsm.copy_(pack_fsm(cur["_explore_state"], cur["_explore_steps"], cur["_explore_dir"],
                                 cur["_photo_avoiding"], cur["_photo_steps"], cur["_photo_dir"],
                                 cur["_antiphoto_avoiding"], cur["_antiphoto_steps"], cur["_antiphoto_dir"]))

    def seed(self, seed: int):
        self._seed = int(seed)

    def inject_noise(self, rab_u=None, turn_dur=None, spawn_u=None, yaw_u=None):
        """Parity mode: use these draws for the NEXT step/reset instead of the Philox stream."""
        E, dev = self.num_envs, self.device
        inj = {}
        if rab_u is not None:
            inj["rab_u"] = torch.as_tensor(rab_u, dtype=torch.float32, device=dev).reshape(E, N, N).contiguous()
        if turn_dur is not None:
            inj["turn_dur"] = torch.as_tensor(turn_dur, device=dev).to(torch.int32).reshape(E, N, 3).contiguous()
        if spawn_u is not None:
            inj["spawn_u"] = torch.as_tensor(spawn_u, dtype=torch.float32, device=dev).reshape(-1, E, N, 2).contiguous()
        if yaw_u is not None:
            inj["yaw_u"] = torch.as_tensor(yaw_u, dtype=torch.float32, device=dev).reshape(E, N).contiguous()
        self._injected = inj

    def _noise(self, steps: int = 1) -> SwarmNoise:
        inj, self._injected = self._injected, {}
        self._noise_keepalive = inj
        nz = SwarmNoise()
        for k in ("rab_u", "turn_dur", "spawn_u", "yaw_u"):
            if k in inj:
                setattr(nz, k, inj[k].data_ptr())
        nz.spawn_rounds = inj["spawn_u"].shape[0] if "spawn_u" in inj else 0
        nz.seed = self._seed
        nz.step_counter = self._step_counter
        nz.env_offset = self._env_offset
        if self.job_reset_clock is not None and steps > 0:
            nz.any_reset_mode = 1
            nz.any_reset_bits = self.job_reset_clock.bits(self._step_counter, steps)
        self._step_counter += 1
        return nz

    def attach_job_reset_clock(self, group=None, peers=None):
        """Multi-GPU jobs with de-synchronised episode counters: make ENV:1262's "any env reset -> re-solve all envs"
        job-wide instead of per shard (see sharding.JobResetClock).  Collective: call it on every rank."""
        from .sharding import JobResetClock
        self._check_len_buf()
        self.job_reset_clock = JobResetClock(self, group, peers)
        return self.job_reset_clock

    def _sync_flags(self):
        """Rebuild the kernel's rotating any-reset flags after episode_length_buf was written from outside."""
        with self._device_guard():
            rc = self._lib.swarm_sync_episode_flags(C.byref(self.params), C.byref(self._state), self._step_counter,
                                                    self.num_envs, self._stream())
        _lib.check(rc, "swarm_sync_episode_flags")
        self._len_version = self.episode_length_buf._version
        if self.job_reset_clock is not None:
            self.job_reset_clock.rebuild(self)

    def _check_len_buf(self):
        if self.episode_length_buf._version != self._len_version:
            self._sync_flags()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self, seed: int | None = None, options: dict | None = None):
        if seed is not None:
            self._seed = int(seed)
        nz = self._noise(0)
        with self._device_guard():
            rc = self._lib.swarm_reset(C.byref(self.params), C.byref(self._state), C.byref(nz),
                                       C.byref(self._critic_out(self.fused_critic)), self.num_envs, self._stream())
        _lib.check(rc, "swarm_reset")
        if self.job_reset_clock is not None:
            self.job_reset_clock.rebuild(self)
        return dict(self._obs_views), self.extras

    def _gather_actions(self, actions) -> torch.Tensor:
        """Return an (E,N,act) contiguous tensor of the right dtype, zero-copy when ``actions`` is
        the usual dict of 20 strided views of one tensor (agents/poca_trainer.py:559)."""
        want = torch.long if self.params.discrete_actions else torch.float32
        E, A = self.num_envs, self.act_dim
        if isinstance(actions, torch.Tensor):
            t = actions.reshape(E, N, A)
            if t.dtype != want or not t.is_contiguous() or t.device != self.device:
                self._act_buf.copy_(t)
                t = self._act_buf
            return t
        a0 = actions[self.possible_agents[0]]
        if (a0.dtype == want and a0.device == self.device and tuple(a0.shape) == (E, A)
                and a0.stride() == (N * A, 1)):
            base, isz = a0.data_ptr(), a0.element_size()
            if all(actions[a].data_ptr() == base + i * A * isz and actions[a].stride() == (N * A, 1)
                   and actions[a].dtype == want for i, a in enumerate(self.possible_agents)):
                # read during the call; callers may mutate their views afterwards (scripts/play.py:702)
                return torch.as_strided(a0, (E, N, A), (N * A, A, 1))
        for i, a in enumerate(self.possible_agents):
            self._act_buf[:, i] = actions[a].reshape(E, A)
        return self._act_buf

    def step_tensor(self, actions: torch.Tensor):
        """One env.step from an (E,N,act) action tensor; returns (obs (E,N,obs), reward (E), time_out (E) bool)
        views of buffers that the next step overwrites."""
        act = self._gather_actions(actions)
        self._check_len_buf()
        nz = self._noise()
        # the critic state rides on the step that the learned cadence says precedes the next get_critic_state()
        want = self.fused_critic and self._steps_since_critic + 1 == self._critic_period
        with self._device_guard():
            rc = self._lib.swarm_step(C.byref(self.params), C.byref(self._state), C.c_void_p(act.data_ptr()),
                                      C.byref(nz), C.byref(self._critic_out(want)), self.num_envs, self._stream())
        _lib.check(rc, "swarm_step")
        self._steps_since_critic += 1
        return self._obs, self._reward, self._time_out.view(torch.bool)

    def step(self, actions: dict):
        _, reward, time_out = self.step_tensor(actions)
        agents = self.possible_agents
        reward_dict = dict.fromkeys(agents, reward)
        terminated = dict.fromkeys(agents, self._terminated)
        truncated = dict.fromkeys(agents, time_out)
        return dict(self._obs_views), reward_dict, terminated, truncated, self.extras

    def rollout(self, actions: torch.Tensor, steps: int | None = None):
        """``steps`` consecutive env.steps with device-resident actions.  ``actions`` is either one action per step,
        shape (T,E,N,act) (then ``steps`` must be None or T), or ONE action (E,N,act) / (E,N) that is repeated
        ``steps`` times (the trainers' decision_period loop; ``steps`` is required).  Returns (last obs, summed
        reward, OR-ed time_out).  Raises ValueError on any other shape: the kernel cannot bounds-check the buffer."""
        want = torch.long if self.params.discrete_actions else torch.float32
        E, A = self.num_envs, self.act_dim
        shape = tuple(actions.shape)
        if shape == (E, N, A) or (A == 1 and shape == (E, N)):
            if steps is None:
                raise ValueError("rollout: `steps` is required when one action is repeated")
            T, stride = int(steps), 0
        elif (len(shape) == 4 and shape[1:] == (E, N, A)) or (A == 1 and len(shape) == 3 and shape[1:] == (E, N)):
            T, stride = int(shape[0]), E * N * A
            if steps is not None and int(steps) != T:
                raise ValueError(f"rollout: steps={steps} does not match the {T} per-step actions given")
        else:
            raise ValueError(f"rollout: actions must be ({E}, {N}, {A}) (one action, repeated `steps` times) or "
                             f"(T, {E}, {N}, {A}) (one per step), got {shape}")
        if T <= 0:
            raise ValueError("rollout: steps must be > 0")
        if self.job_reset_clock is not None and T > 32:
            raise ValueError("rollout: with a job-wide reset clock attached one call covers at most 32 steps "
                             "(SwarmNoise.any_reset_bits is a 32-bit schedule)")
        if actions.dtype != want or actions.device != self.device:
            actions = actions.to(device=self.device, dtype=want)
        actions = actions.contiguous()
        self._check_len_buf()
        nz = self._noise(T)
        self._step_counter += T - 1
        with self._device_guard():
            rc = self._lib.swarm_rollout(C.byref(self.params), C.byref(self._state), C.c_void_p(actions.data_ptr()),
                                         stride, C.byref(nz), C.byref(self._critic_out(self.fused_critic)), E, T,
                                         self._stream())
        _lib.check(rc, "swarm_rollout")
        self._steps_since_critic += T
        return self._obs, self._reward, self._time_out.view(torch.bool)

    def step_host(self, actions_host: torch.Tensor, obs_host: torch.Tensor, reward_host: torch.Tensor,
                  time_out_host: torch.Tensor):
        """One env.step for a caller whose buffers live in (pinned) HOST memory: ``swarm_host_step`` uploads the
        (E,N,act) action batch, steps, and downloads obs (E,N,obs) / reward (E) / time_out (E) uint8 into the given
        host tensors, chunked so the PCIe copies overlap the kernel.  Returns after the results have landed."""
        want = torch.long if self.params.discrete_actions else torch.float32
        E = self.num_envs
        for name, t, dt, numel in (("actions_host", actions_host, want, E * N * self.act_dim),
                                   ("obs_host", obs_host, torch.float32, E * N * self.obs_dim),
                                   ("reward_host", reward_host, torch.float32, E),
                                   ("time_out_host", time_out_host, torch.uint8, E)):
            if t.device.type != "cpu" or t.dtype != dt or t.numel() != numel or not t.is_contiguous():
                raise ValueError(f"step_host: {name} must be a contiguous CPU {dt} tensor with {numel} elements")
        self._check_len_buf()
        nz = self._noise()
        with self._device_guard():
            rc = self._lib.swarm_host_step(C.byref(self.params), C.byref(self._state), C.c_void_p(actions_host.data_ptr()),
                                           C.byref(nz), C.c_void_p(obs_host.data_ptr()), C.c_void_p(reward_host.data_ptr()),
                                           C.c_void_p(time_out_host.data_ptr()), C.c_void_p(self._act_buf.data_ptr()),
                                           C.byref(self._critic_out(False)), E, self._stream())
        _lib.check(rc, "swarm_host_step")
        self._steps_since_critic += 1
        return obs_host, reward_host, time_out_host

    def _critic_out(self, want: bool) -> SwarmOut:
        """The SwarmOut block of the next call: with ``want`` the kernel also writes get_critic_state() of the state
        it leaves behind into the current pool buffer."""
        self._critic_fresh = bool(want)
        self._critic_last = None           # the state is about to change: the tensor handed out last is history
        if not want:
            return self._out
        self._out_c.critic = self._critic_pool[self._critic_slot].data_ptr()
        self._pose_version = (self._agent_pos._version, self._agent_yaw._version)
        return self._out_c

    def get_critic_state(self) -> torch.Tensor:
        """(E,N,5) = (rho, cos alpha, sin alpha, cos beta, sin beta), ENV:1279-1290.

        With ``fused_critic`` (default) the step kernel computes it in its epilogue, while the pose is still in
        registers: reset() and rollout() always do, step() does on the step that - by the cadence observed so far
        (the trainers ask once per decision, every ``decision_period`` steps) - precedes the next call.  This call
        then launches nothing and returns one of CRITIC_POOL rotating buffers: the tensor stays untouched until
        CRITIC_POOL - 1 further get_critic_state() calls (the reference's trainers copy it into their rollout buffer
        within the same decision).  When the prediction missed, or agent_pos / agent_yaw were written from outside
        since, it falls back to the stand-alone kernel.  ``fused_critic=False`` returns a fresh tensor per call."""
        if not self.fused_critic:
            out = torch.empty(self.num_envs, N, 5, dtype=torch.float32, device=self.device)
        else:
            pose_now = (self._agent_pos._version, self._agent_yaw._version)
            if self._critic_last is not None and self._pose_version == pose_now:
                return self._critic_last   # asked again for the same state (OC2: next_critic_state, then critic_state)
            out = self._critic_pool[self._critic_slot]
            self._critic_slot = (self._critic_slot + 1) % self.CRITIC_POOL
            if self._steps_since_critic > 0:
                self._critic_period = self._steps_since_critic
            self._steps_since_critic = 0
            fresh = self._critic_fresh and self._pose_version == pose_now
            self._critic_fresh = False
            self._critic_last = out
            if fresh:
                return out
            self._pose_version = pose_now
        with self._device_guard():
            rc = self._lib.swarm_critic_state(C.byref(self.params), C.byref(self._state), C.c_void_p(out.data_ptr()),
                                              self.num_envs, self._stream())
        _lib.check(rc, "swarm_critic_state")
        return out

    def close(self):
        pass

    # ── checkpoint / resume (the reference never saves env state) ───────────
    _STATE_TENSORS = {
        "pos": "agent_pos", "yaw": "agent_yaw", "prev_ground": "prev_ground_color", "cached_left": "_cached_left_vel",
        "cached_right": "_cached_right_vel", "fsm": "_fsm", "beh_cache": "_beh_cache", "mission_flags": "_mission_flags",
        "episode_length_buf": "episode_length_buf", "episode_group_reward": "_episode_group_reward",
        "completed_group_reward": "completed_group_reward",
        "completed_terminal_critic_state": "completed_terminal_critic_state", "obs": "_obs",
    }

    def state_dict(self) -> dict:
        """Everything needed to resume a run bit for bit: the SwarmState arrays, the last observation and the Philox
        position (seed, step counter, env offset).  Tensors are detached CPU copies (``torch.save``-able)."""
        sd = {k: getattr(self, attr).detach().cpu().clone() for k, attr in self._STATE_TENSORS.items()}
        sd["meta"] = {"seed": self._seed, "step_counter": self._step_counter, "env_offset": self._env_offset,
                      "num_envs": self.num_envs, "mission": int(self.params.mission), "obs_dim": self.obs_dim,
                      "discrete_actions": bool(self.params.discrete_actions)}
        return sd

    def load_state_dict(self, sd: dict):
        meta = sd["meta"]
        mine = (self.num_envs, int(self.params.mission), self.obs_dim, bool(self.params.discrete_actions))
        theirs = (meta["num_envs"], meta["mission"], meta["obs_dim"], meta["discrete_actions"])
        if mine != theirs:
            raise ValueError(f"checkpoint is for (envs, mission, obs_dim, discrete) = {theirs}, this env is {mine}")
        for k, attr in self._STATE_TENSORS.items():
            dst = getattr(self, attr)
            dst.copy_(sd[k].to(dst.dtype).reshape(dst.shape))
        self._seed, self._step_counter = int(meta["seed"]), int(meta["step_counter"])
        self._env_offset = int(meta["env_offset"])
        self._injected = {}
        self._critic_fresh = False
        self._sync_flags()

    # ── teacher-forcing helpers for the parity tests ────────────────────────────────────────
    def load_state(self, state: dict):
        """Overwrite the device state from host arrays in the include/swarm_abi.h layouts."""
        m = {
            "pos": self.agent_pos, "yaw": self.agent_yaw, "prev_ground": self.prev_ground_color,
            "cached_left": self._cached_left_vel, "cached_right": self._cached_right_vel, "fsm": self._fsm,
            "beh_cache": self._beh_cache, "mission_flags": self._mission_flags,
            "episode_length_buf": self.episode_length_buf, "episode_group_reward": self._episode_group_reward,
            "completed_group_reward": self.completed_group_reward,
            "completed_terminal_critic_state": self.completed_terminal_critic_state,
        }
        for k, dst in m.items():
            dst.copy_(torch.as_tensor(state[k]).to(dst.dtype).reshape(dst.shape))
        self._critic_fresh = False
        self._sync_flags()

    def dump_state(self) -> dict:
        m = {
            "pos": self.agent_pos, "yaw": self.agent_yaw, "prev_ground": self.prev_ground_color,
            "cached_left": self._cached_left_vel, "cached_right": self._cached_right_vel, "fsm": self._fsm,
            "beh_cache": self._beh_cache, "mission_flags": self._mission_flags,
            "episode_length_buf": self.episode_length_buf, "episode_group_reward": self._episode_group_reward,
            "completed_group_reward": self.completed_group_reward,
            "completed_terminal_critic_state": self.completed_terminal_critic_state,
        }
        out = {k: v.detach().cpu().numpy().copy() for k, v in m.items()}
        out["fsm"] &= FSM_STATE_MASK   # bits 18..23 hold pre-drawn turn-duration bits, not reference state
        return out


# ── registry (missions/*/__init__.py of the reference) ─────────────────────────────────────────
registry = {
    task_id: {
        "entry_point": f"{__name__}:SwarmEnv",
        "disable_env_checker": True,
        "kwargs": {"env_cfg_entry_point": f"{cfg_cls.__module__}:{cfg_cls.__name__}"},
    }
    for task_id, cfg_cls in TASK_CFGS.items()
}


def make(task_id: str, cfg=None, **kwargs) -> SwarmEnv:
    """``gym.make(task_id, cfg=cfg)`` equivalent that needs no gymnasium install."""
    if task_id not in TASK_CFGS:
        raise KeyError(f"unknown task {task_id!r}; known: {sorted(TASK_CFGS)}")
    if cfg is None:
        cfg = TASK_CFGS[task_id]()
    elif not isinstance(cfg, TASK_CFGS[task_id]):
        raise TypeError(f"{task_id} expects a {TASK_CFGS[task_id].__name__}, got {type(cfg).__name__}")
    return SwarmEnv(cfg, **kwargs)


def register_gym() -> bool:
    """Register the seven ids with gymnasium when it is installed; returns False otherwise."""
    try:
        import gymnasium as gym
    except ImportError:
        return False
    for task_id, spec in registry.items():
        if task_id not in gym.registry:
            gym.register(id=task_id, entry_point=spec["entry_point"], disable_env_checker=True,
                         kwargs=dict(spec["kwargs"]))
    return True
