"""Loader of the C-ABI library (libswarmstep.so).  There is no CPU or eager-torch fallback: if the
library cannot be built/loaded the import of the step path fails loudly."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build
from .params import ABI_VERSION, SwarmNoise, SwarmOut, SwarmParams, SwarmState

_lib = None

EXPORTS = (
    "swarm_step", "swarm_reset", "swarm_critic_state", "swarm_rollout", "swarm_host_step",
    "swarm_abi_version", "swarm_kernel_launch_count", "swarm_last_error_string", "swarm_fp32_peak",
    "swarm_detmath_eval", "swarm_sync_episode_flags", "swarm_mc_tick", "swarm_mc_reset", "swarm_host_release",
)


class SwarmLibraryError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """dlopen libswarmstep.so (building it in-tree first if nvcc is available and it is stale)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("SWARM_LIB_OVERRIDE") or _build.LIB  # override = tuning variants only
    if build_if_missing and path == _build.LIB and _build.needs_build():
        try:
            _build.build()
        except Exception as exc:  # stale-but-present library is still usable on a box without nvcc
            if not os.path.exists(path):
                raise SwarmLibraryError(f"libswarmstep.so is missing and could not be built: {exc}") from exc
            if os.environ.get("SWARM_STRICT_BUILD"):
                raise SwarmLibraryError(f"libswarmstep.so is older than its sources and the rebuild failed: {exc}") from exc
            import warnings
            warnings.warn(f"libswarmstep.so is OLDER than csrc/swarm_step.cu / swarm_abi.h and could not be rebuilt "
                          f"({str(exc)[:200]}); loading the stale library (set SWARM_STRICT_BUILD=1 to make this an "
                          f"error)", RuntimeWarning, stacklevel=2)
    if not os.path.exists(path):
        raise SwarmLibraryError(f"{path} not found; run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(path)
    P, S, Nz, O = C.POINTER(SwarmParams), C.POINTER(SwarmState), C.POINTER(SwarmNoise), C.POINTER(SwarmOut)
    lib.swarm_step.argtypes = [P, S, C.c_void_p, Nz, O, C.c_int, C.c_void_p]
    lib.swarm_reset.argtypes = [P, S, Nz, O, C.c_int, C.c_void_p]
    lib.swarm_critic_state.argtypes = [P, S, C.c_void_p, C.c_int, C.c_void_p]
    lib.swarm_sync_episode_flags.argtypes = [P, S, C.c_uint64, C.c_int, C.c_void_p]
    lib.swarm_rollout.argtypes = [P, S, C.c_void_p, C.c_int64, Nz, O, C.c_int, C.c_int, C.c_void_p]
    lib.swarm_host_step.argtypes = [P, S, C.c_void_p, Nz, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, O,
                                    C.c_int, C.c_void_p]
    lib.swarm_fp32_peak.argtypes = [C.c_int, C.POINTER(C.c_float), C.c_void_p]
    lib.swarm_mc_tick.argtypes = [P, S, C.c_void_p, C.c_void_p, Nz, O, C.c_int, C.c_int, C.c_void_p]
    lib.swarm_mc_reset.argtypes = [P, S, Nz, C.c_int, C.c_void_p]
    lib.swarm_detmath_eval.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_void_p]
    for name in EXPORTS:
        getattr(lib, name).restype = C.c_int
    lib.swarm_last_error_string.restype = C.c_char_p
    if lib.swarm_abi_version() != ABI_VERSION:
        raise SwarmLibraryError("libswarmstep.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str):
    """Map the C status to the reference-style Python exceptions (Errors)."""
    if rc == 0:
        return
    msg = load().swarm_last_error_string().decode()
    if rc < 0:
        raise ValueError(f"{what}: bad argument ({rc}): {msg}")
    raise RuntimeError(f"{what}: CUDA error {rc}: {msg}")
