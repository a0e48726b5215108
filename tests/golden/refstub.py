"""Stub harness that lets the UNMODIFIED reference env classes run on CPU in the build container.

Test infrastructure only (fixture generation).  Nothing here is imported by the product
package, by the ``-m gpu`` tests, by ``smoke()`` or by ``bench.py``: ``/root/reference`` does
not exist on the GPU box.  The stubs restate the un-vendored ``isaaclab`` orchestration the
reference relies on:

* ``isaaclab.utils.configclass``  - class decorator whose ``__init__`` deep-copies class-level
  defaults into the instance (reference cfg classes use mutable class defaults,
  directional_gate_env_cfg.py:85-110).
* ``isaaclab.envs.DirectMARLEnv.step``  - hook order pre_physics -> apply_action x decimation ->
  episode_length_buf += 1 -> get_dones -> get_rewards -> reset_idx(timed-out) ->
  get_observations (directional_gate_env.py:66-68, :1202; homing_env.py:88 rely on it).
* inert ``isaaclab.sim`` / ``isaaclab.markers`` / ``omni`` / ``pxr`` / ``gymnasium.register``.
"""
from __future__ import annotations

import copy
import importlib
import math
import sys
import types

import torch

import os

REF_ROOT = "/root/reference"
_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def find_reference_package():
    """Directory of the reference's ``SwarmACB_isaac`` python package: the read-only checkout in the build container,
    or the offline pip install under baseline/_ref (tools/install_reference.sh; that one travels to the GPU box)."""
    for cand in (os.environ.get("SWARM_REFERENCE_PKG"), REF_ROOT + "/source/SwarmACB_isaac/SwarmACB_isaac",
                 os.path.join(_REPO, "baseline", "_ref", "SwarmACB_isaac")):
        if cand and os.path.isdir(os.path.join(cand, "tasks", "direct", "epuck")):
            return cand
    return None


_PKG = find_reference_package() or REF_ROOT + "/source/SwarmACB_isaac/SwarmACB_isaac"

MISSION_MODULES = {
    "dgt": ("directional_gate", "DirectionalGateEnv", "DirectionalGateEnvCfg"),
    "xor": ("xor_aggregation", "XorAggregationEnv", "XorAggregationEnvCfg"),
    "hom": ("homing", "HomingEnv", "HomingEnvCfg"),
    "for": ("foraging", "ForagingEnv", "ForagingEnvCfg"),
    "shl": ("sheltering", "ShelteringEnv", "ShelteringEnvCfg"),
}


def _configclass(cls):
    def __init__(self, **kwargs):
        seen = set()
        for klass in type(self).__mro__:
            for name, value in vars(klass).items():
                if name.startswith("__") or name in seen:
                    continue
                seen.add(name)
                if callable(value) or isinstance(value, (staticmethod, classmethod, property)):
                    continue
                setattr(self, name, copy.deepcopy(value))
        for k, v in kwargs.items():
            setattr(self, k, v)

    cls.__init__ = __init__
    return cls


class _Inert:
    def __init__(self, *a, **k):
        self.__dict__.update(k)

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return _Inert()


class _SimulationCfg:
    def __init__(self, dt=0.1, render_interval=1, gravity=(0.0, 0.0, -9.81), device="cpu"):
        self.dt, self.render_interval, self.gravity, self.device = dt, render_interval, gravity, device


class _SceneCfg:
    def __init__(self, num_envs=1, env_spacing=4.0, replicate_physics=True):
        self.num_envs, self.env_spacing, self.replicate_physics = num_envs, env_spacing, replicate_physics


@_configclass
class _DirectMARLEnvCfg:
    seed = None
    decimation = 1
    episode_length_s = 1.0


class _Sim:
    def has_gui(self):
        return False


class _DirectMARLEnv:
    """Restatement of isaaclab.envs.DirectMARLEnv orchestration (no simulator)."""

    def __init__(self, cfg, render_mode=None, **kwargs):
        self.cfg = cfg
        self.device = torch.device(cfg.sim.device)
        self.num_envs = cfg.scene.num_envs
        self.scene = types.SimpleNamespace(num_envs=self.num_envs)
        self.sim = _Sim()
        self.max_episode_length = math.ceil(cfg.episode_length_s / (cfg.sim.dt * cfg.decimation))
        self.episode_length_buf = torch.zeros(self.num_envs, dtype=torch.long, device=self.device)
        self.reset_buf = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        self.extras = {}

    @property
    def unwrapped(self):
        return self

    def _reset_idx(self, env_ids):
        self.episode_length_buf[env_ids] = 0

    def reset(self, seed=None, options=None):
        self._reset_idx(torch.arange(self.num_envs, device=self.device))
        return self._get_observations(), self.extras

    def step(self, actions):
        self._pre_physics_step(actions)
        for _ in range(self.cfg.decimation):
            self._apply_action()
        self.episode_length_buf += 1
        terminated, time_outs = self._get_dones()
        a0 = self.cfg.possible_agents[0]
        self.reset_buf[:] = terminated[a0] | time_outs[a0]
        rewards = self._get_rewards()
        reset_ids = self.reset_buf.nonzero(as_tuple=False).squeeze(-1)
        if len(reset_ids) > 0:
            self._reset_idx(reset_ids)
        obs = self._get_observations()
        return obs, rewards, terminated, time_outs, self.extras

    def close(self):
        pass


def install():
    """Install the stub modules (idempotent)."""
    if "isaaclab" in sys.modules and getattr(sys.modules["isaaclab"], "_swarm_stub", False):
        return
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    il = mod("isaaclab", _swarm_stub=True)
    il.utils = mod("isaaclab.utils", configclass=_configclass)
    il.sim = mod("isaaclab.sim", SimulationCfg=_SimulationCfg, DomeLightCfg=_Inert, CuboidCfg=_Inert,
                 SphereCfg=_Inert, CylinderCfg=_Inert, PreviewSurfaceCfg=_Inert)
    il.scene = mod("isaaclab.scene", InteractiveSceneCfg=_SceneCfg)
    il.markers = mod("isaaclab.markers", VisualizationMarkersCfg=_Inert, VisualizationMarkers=_Inert)
    il.envs = mod("isaaclab.envs", DirectMARLEnv=_DirectMARLEnv, DirectMARLEnvCfg=_DirectMARLEnvCfg)
    omni = mod("omni")
    omni.usd = mod("omni.usd", get_context=_Inert())
    mod("pxr", Gf=_Inert(), UsdGeom=_Inert(), Vt=_Inert())
    registry = {}
    mod("gymnasium", register=lambda id, **kw: registry.__setitem__(id, kw), registry=registry)

    def ns(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m

    ns("SwarmACB_isaac", _PKG)
    ns("SwarmACB_isaac.tasks", _PKG + "/tasks")
    ns("SwarmACB_isaac.tasks.direct", _PKG + "/tasks/direct")


def load_mission(mission: str):
    """Return (EnvClass, CfgClass) of the reference for a mission key in MISSION_MODULES."""
    install()
    pkg, env_name, cfg_name = MISSION_MODULES[mission]
    base = f"SwarmACB_isaac.tasks.direct.missions.{pkg}"
    env_mod = importlib.import_module(f"{base}.{pkg}_env")
    cfg_mod = importlib.import_module(f"{base}.{pkg}_env_cfg")
    return getattr(env_mod, env_name), getattr(cfg_mod, cfg_name)


def make_ref_env(mission: str, mode: str, num_envs: int, decimation: int = 1, **overrides):
    """Build a reference env the way scripts/train.py:166-188 does.

    mode: "cyclamen"/"lily"/"tulip"/"daisy"/"dandelion" -> update_variant(mode);
          "oc2"   -> cyclamen + use_continuous_actions(full_observations=True);
          "oc2c"  -> cyclamen + use_continuous_actions(full_observations=False).
    """
    Env, Cfg = load_mission(mission)
    cfg = Cfg()
    if mode in ("oc2", "oc2c"):
        cfg.update_variant("cyclamen")
        cfg.use_continuous_actions(full_observations=(mode == "oc2"))
    else:
        cfg.update_variant(mode)
    cfg.scene.num_envs = num_envs
    cfg.decimation = decimation
    for k, v in overrides.items():
        assert hasattr(cfg, k), k
        setattr(cfg, k, v)
    return Env(cfg)


class NoiseTap:
    """Context manager recording every torch.rand / torch.randint draw the reference makes."""

    def __init__(self):
        self.rand_calls = []
        self.randint_calls = []

    def __enter__(self):
        self._rand, self._randint = torch.rand, torch.randint
        tap = self

        def rand(*size, **kw):
            out = tap._rand(*size, **kw)
            tap.rand_calls.append(out.clone())
            return out

        def randint(*a, **kw):
            out = tap._randint(*a, **kw)
            tap.randint_calls.append(out.clone())
            return out

        torch.rand, torch.randint = rand, randint
        return self

    def __exit__(self, *exc):
        torch.rand, torch.randint = self._rand, self._randint
        return False
