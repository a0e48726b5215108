"""swarmacb-isaaclab_b200: B200-native fused swarm step behind the SwarmACB-isaaclab task API.

Only the hot path lives here: the host-side mirror of the reference's env/cfg
interface and the CUDA kernels + C ABI under ``csrc/``.  Import as ``swarmacb_isaaclab_b200``.
"""
from .cfg import (  # noqa: F401
    DirectionalGateEnvCfg, XorAggregationEnvCfg, HomingEnvCfg, ForagingEnvCfg, ShelteringEnvCfg,
    TASK_CFGS, MISSION_CFGS,
)
from .params import build_params, SwarmParams  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # env / registry import torch-side machinery lazily so cfg/params stay importable anywhere
    if name in ("SwarmEnv", "make", "register_gym", "registry"):
        from . import env as _env
        return getattr(_env, name)
    raise AttributeError(name)
