"""TEST INFRASTRUCTURE: a stand-in for the part of ``gymnasium`` the reference uses to expose its tasks
(missions/*/__init__.py: ``gym.register(id=..., entry_point="module:Class", disable_env_checker=True,
kwargs={"env_cfg_entry_point": "module:Class"})``; scripts/train.py:96-106,188: ``gym.spec(id).kwargs`` and
``gym.make(id, cfg=cfg)``).  gymnasium itself is not installed in this image (no network); the stand-in follows its
documented registration semantics: string entry points are resolved with importlib, ``make`` merges the spec's kwargs
with the call's and instantiates the entry point."""
import importlib
import sys
import types


class EnvSpec:
    def __init__(self, id, entry_point, kwargs, **extra):
        self.id, self.entry_point, self.kwargs, self.extra = id, entry_point, dict(kwargs or {}), extra


def install():
    """Put the stand-in into sys.modules as ``gymnasium`` (returns the module; idempotent)."""
    if "gymnasium" in sys.modules and getattr(sys.modules["gymnasium"], "_swarm_test_stub", False):
        return sys.modules["gymnasium"]
    gym = types.ModuleType("gymnasium")
    gym._swarm_test_stub = True
    gym.registry = {}

    def register(id, entry_point=None, kwargs=None, **extra):
        gym.registry[id] = EnvSpec(id, entry_point, kwargs, **extra)

    def spec(id):
        return gym.registry[id]

    def make(id, **kwargs):
        s = gym.registry[id]
        ep = s.entry_point
        if isinstance(ep, str):
            mod, _, name = ep.partition(":")
            ep = getattr(importlib.import_module(mod), name)
        return ep(**{**s.kwargs, **kwargs})

    class Box:
        def __init__(self, low, high, shape, dtype):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    class Discrete:
        def __init__(self, n):
            self.n = int(n)

    gym.register, gym.spec, gym.make = register, spec, make
    gym.spaces = types.ModuleType("gymnasium.spaces")
    gym.spaces.Box, gym.spaces.Discrete = Box, Discrete
    sys.modules["gymnasium"] = gym
    sys.modules["gymnasium.spaces"] = gym.spaces
    return gym


def uninstall():
    for k in ("gymnasium", "gymnasium.spaces"):
        if getattr(sys.modules.get(k), "_swarm_test_stub", False) or k == "gymnasium.spaces":
            sys.modules.pop(k, None)
