"""TEST INFRASTRUCTURE: ``SwarmEnv``'s host logic on CPU tensors with the C oracle standing in for libswarmstep.so.

Lets the build container (no GPU) exercise everything *above* the C ABI - the dict protocol, attribute
surface, zero-copy action gathering, counters, auto-reset semantics - against the reference's own trainers
(tests/test_trainer_protocol.py).  It is not a product path: ``SwarmEnv`` itself refuses any non-CUDA device.
"""
from __future__ import annotations

import contextlib
import ctypes as C

import numpy as np
import torch

from oracle import oracle
from swarmacb_isaaclab_b200.env import SwarmEnv
from swarmacb_isaaclab_b200.params import N


def _obj(ref):
    return ref._obj if hasattr(ref, "_obj") else ref


def _np(ptr, n, ctype):
    addr = ptr if isinstance(ptr, int) else C.cast(ptr, C.c_void_p).value
    return np.ctypeslib.as_array((ctype * n).from_address(addr))


class OracleLib:
    """Same entry points and argument lists as the ctypes handle of libswarmstep.so (stream ignored); noise that
    the kernel would draw from Philox is drawn from numpy and injected into the oracle."""

    def __init__(self, seed: int = 0):
        self.rng = np.random.default_rng(seed)
        self.lib = oracle.lib()
        self.steps = 0

    def _draw(self, p, nz, E, keep):
        rng = self.rng
        if not nz.rab_u:
            a = rng.random((E, N, N), dtype=np.float32); keep.append(a); nz.rab_u = a.ctypes.data
        if p.discrete_actions and not nz.turn_dur:
            a = rng.integers(1, 5, (E, N, 3)).astype(np.int32); keep.append(a); nz.turn_dur = a.ctypes.data
        if not nz.spawn_u:
            a = rng.random((8, E, N, 2), dtype=np.float32); keep.append(a); nz.spawn_u = a.ctypes.data
            nz.spawn_rounds = 8
        if not nz.yaw_u:
            a = rng.random((E, N), dtype=np.float32); keep.append(a); nz.yaw_u = a.ctypes.data

    def swarm_step(self, p, st, actions, nz, out, E, stream):
        p, st, nz, out = _obj(p), _obj(st), _obj(nz), _obj(out)
        keep = []
        self._draw(p, nz, E, keep)
        self.steps += 1
        return self.lib.swarm_oracle_step(C.byref(p), C.byref(st), actions, C.byref(nz), C.byref(out), E)

    def swarm_reset(self, p, st, nz, out, E, stream):
        p, st, nz, out = _obj(p), _obj(st), _obj(nz), _obj(out)
        keep = []
        self._draw(p, nz, E, keep)
        return self.lib.swarm_oracle_reset(C.byref(p), C.byref(st), C.byref(nz), C.byref(out), E)

    def swarm_rollout(self, p, st, actions, stride, nz, out, E, T, stream):
        p_, out_ = _obj(p), _obj(out)
        elem = 8 if p_.discrete_actions else 4
        reward = _np(out_.reward, E, C.c_float)
        time_out = _np(out_.time_out, E, C.c_uint8)
        acc_r, acc_t = np.zeros(E, np.float32), np.zeros(E, np.uint8)
        base = actions.value if isinstance(actions, C.c_void_p) else int(actions)
        for t in range(T):
            nz_t = type(_obj(nz))()
            C.memmove(C.byref(nz_t), C.byref(_obj(nz)), C.sizeof(nz_t))
            rc = self.swarm_step(p, st, C.c_void_p(base + t * int(stride) * elem), nz_t, out, E, stream)
            if rc:
                return rc
            acc_r += reward
            acc_t |= time_out
        reward[:] = acc_r
        time_out[:] = acc_t
        return 0

    def swarm_critic_state(self, p, st, critic_out, E, stream):
        return self.lib.swarm_oracle_critic_state(C.byref(_obj(p)), C.byref(_obj(st)), critic_out, E)

    def swarm_sync_episode_flags(self, p, st, next_counter, E, stream):
        return 0

    def swarm_last_error_string(self):
        return b"oracle"


class OracleBackedEnv(SwarmEnv):
    """``SwarmEnv`` with CPU tensors and the oracle behind the C-ABI call sites (tests only)."""

    def _resolve_device(self, name):
        return torch.device("cpu")

    def _load_library(self):
        return OracleLib(seed=int(getattr(self.cfg, "seed", 0) or 0))

    def _device_guard(self):
        return contextlib.nullcontext()

    def _stream(self):
        return None
