"""include/swarm_detmath.h: accuracy against double libm (CPU) and CPU == GPU bit for bit (gpu)."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import oracle


def _ulp_err(got, exact):
    got = np.asarray(got, np.float64)
    ulp = np.spacing(np.abs(exact).astype(np.float32)).astype(np.float64)
    return np.abs(got - exact) / ulp


def _inputs(n=400_000, seed=0):
    rng = np.random.default_rng(seed)
    a = (rng.random(n, dtype=np.float32) * 2 - 1) * np.float32(2 * math.pi)
    a[:1000] = np.linspace(-7, 7, 1000, dtype=np.float32)
    a[1000:1010] = [0.0, -0.0, math.pi, -math.pi, math.pi / 2, -math.pi / 2, 1e-8, -1e-8, 3.1415925, 6.2831855]
    y = rng.standard_normal(n).astype(np.float32) * np.float32(10.0) ** rng.integers(-6, 2, n).astype(np.float32)
    x = rng.standard_normal(n).astype(np.float32) * np.float32(10.0) ** rng.integers(-6, 2, n).astype(np.float32)
    return a, y, x


def test_sincos_accuracy():
    a, _, _ = _inputs()
    sn, cs = oracle.detmath_sincos(a)
    assert _ulp_err(sn, np.sin(a.astype(np.float64))).max() < 2.0
    assert _ulp_err(cs, np.cos(a.astype(np.float64))).max() < 2.0
    # the vast majority of results are the correctly rounded value
    assert (sn == np.sin(a.astype(np.float64)).astype(np.float32)).mean() > 0.80


def test_atan2_accuracy_and_conventions():
    _, y, x = _inputs()
    out = oracle.detmath_atan2(y, x)
    exact = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    assert _ulp_err(out, exact).max() < 2.5
    assert np.all(np.abs(out) <= np.float32(math.pi))
    z, nz_ = np.float32(0.0), np.float32(-0.0)
    pi = np.float32(math.pi)
    cases = [((z, z), z), ((nz_, z), nz_), ((z, nz_), pi), ((nz_, nz_), -pi), ((z, np.float32(1)), z),
             ((z, np.float32(-1)), pi), ((nz_, np.float32(-1)), -pi), ((np.float32(1), z), np.float32(math.pi / 2)),
             ((np.float32(-1), z), np.float32(-math.pi / 2))]
    for (yy, xx), want in cases:
        got = oracle.detmath_atan2(np.array([yy]), np.array([xx]))[0]
        assert got == want and np.signbit(got) == np.signbit(want), (yy, xx, got, want)


def test_prox_angle_boundary_matches_reference():
    """BEH:251 `abs(prox_angle) <= pi/2` for a lone hit on the 90-degree sensor: the reference's atan2 returns
    exactly -float32(pi/2) there (probed on the reference's torch build), so the obstacle counts as in front."""
    from swarmacb_isaaclab_b200.params import sensor_tables
    cos_a, sin_a, _, _ = sensor_tables()
    v = np.linspace(0.01, 1.0, 5000, dtype=np.float32)
    ang = oracle.detmath_atan2(v * sin_a[2], v * cos_a[2])
    assert np.all(ang == -np.float32(math.pi / 2))
    assert np.all(np.abs(ang) <= np.float32(math.pi * 0.5))


@pytest.mark.gpu
def test_device_results_are_bit_identical_to_host():
    import torch
    from swarmacb_isaaclab_b200 import _lib
    a, y, x = _inputs(seed=1)
    lib = _lib.load()
    outs = []
    for first, second in ((a, a), (y, x)):
        d1, d2 = torch.as_tensor(first, device="cuda:0"), torch.as_tensor(second, device="cuda:0")
        sn, cs, at = torch.empty_like(d1), torch.empty_like(d1), torch.empty_like(d1)
        rc = lib.swarm_detmath_eval(C.c_void_p(d1.data_ptr()), C.c_void_p(d2.data_ptr()), C.c_void_p(sn.data_ptr()),
                                    C.c_void_p(cs.data_ptr()), C.c_void_p(at.data_ptr()), d1.numel(),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "swarm_detmath_eval")
        torch.cuda.synchronize()
        outs.append((sn.cpu().numpy(), cs.cpu().numpy(), at.cpu().numpy()))
    hs, hc = oracle.detmath_sincos(a)
    assert np.array_equal(outs[0][0].view(np.uint32), hs.view(np.uint32))
    assert np.array_equal(outs[0][1].view(np.uint32), hc.view(np.uint32))
    assert np.array_equal(outs[1][2].view(np.uint32), oracle.detmath_atan2(y, x).view(np.uint32))
