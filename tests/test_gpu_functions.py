"""Unit parity of the device functions: the ma_probe_* entry points of
the C ABI against the reference's own functions (tests/golden/unit_functions.npz, made by
tests/golden/make_golden_unit.py from oracle/_ref/unit_oracle).  STRICT arithmetic: bit for bit.  FAST arithmetic
(FMA contraction, Newton reciprocals, the Roe dissipation without tangent / binormal): within 5e-15 of the
magnitude of the terms that are summed — the tolerance is stated per function below."""
import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu

RGAS, GAMMA, CP = 287.05, 1.4, 1004.0


@pytest.fixture(scope="module")
def g():
    return parity.golden("unit_functions")


def _scaled_err(test, ref, scale):
    return float((np.abs(test - ref) / scale).max())


def test_roe_flux(lib, g):
    import miniaero_b200 as ma
    x, ref = g["roe_in"], g["roe_out"]
    vl, vr, a, t, b = x[:, 0:5], x[:, 5:10], x[:, 10:13], x[:, 13:16], x[:, 16:19]
    strict = ma.probe_roe_flux(vl, vr, a, t, b, arith=ma.ARITH_STRICT)
    assert parity.max_ulp(strict, ref) == 0
    fast = ma.probe_roe_flux(vl, vr, a, t, b, arith=ma.ARITH_FAST)
    # magnitude of the summed terms: area * rho * (|u| + c) times 1, (|u| + c), H
    area = np.linalg.norm(a, axis=1)
    rho = np.maximum(vl[:, 0], vr[:, 0])
    T = np.maximum(vl[:, 4], vr[:, 4])
    speed = np.maximum(np.linalg.norm(vl[:, 1:4], axis=1), np.linalg.norm(vr[:, 1:4], axis=1)) + np.sqrt(GAMMA * RGAS * T)
    m = area * rho * speed
    scale = np.column_stack([m, m * speed, m * speed, m * speed, m * (CP * T + 0.5 * speed ** 2)])
    assert _scaled_err(fast, ref, scale) < 5e-15


def test_viscous_flux(lib, g):
    import miniaero_b200 as ma
    x, ref = g["viscous_in"], g["viscous_out"]
    grad, v, a = x[:, 0:15].reshape(-1, 5, 3), x[:, 15:20], x[:, 20:23]
    strict = ma.probe_viscous_flux(grad, v, a, arith=ma.ARITH_STRICT)
    assert parity.max_ulp(strict, ref) == 0
    fast = ma.probe_viscous_flux(grad, v, a, arith=ma.ARITH_FAST)
    T = v[:, 4]
    mu = 1.458e-6 * T * np.sqrt(T) / (T + 110.4)
    area = np.linalg.norm(a, axis=1)
    gmax = np.abs(grad).max(axis=(1, 2))
    mom = 4.0 * mu * gmax * area
    en = mom * np.linalg.norm(v[:, 1:4], axis=1) + mu * (1006.0 / 0.71) * gmax * area
    scale = np.column_stack([np.ones_like(mom), mom, mom, mom, en])
    assert _scaled_err(fast, ref, scale) < 5e-15


def test_primitives(lib, g):
    import miniaero_b200 as ma
    x, ref = g["primitives_in"], g["primitives_out"]
    assert parity.max_ulp(ma.probe_primitives(x, arith=ma.ARITH_STRICT), ref) == 0
    fast = ma.probe_primitives(x, arith=ma.ARITH_FAST)
    # T = (E - k) (gamma - 1) / R is a difference of two energies: scale by the total specific energy
    e_tot = x[:, 4] / x[:, 0] * (GAMMA - 1.0) / RGAS
    scale = np.column_stack([np.abs(ref[:, 0]), np.linalg.norm(ref[:, 1:4], axis=1)[:, None].repeat(3, 1), e_tot])
    assert _scaled_err(fast, ref, scale) < 2e-15


def test_venkat_limiter(lib, g):
    import miniaero_b200 as ma
    x, ref = g["venkat_in"], g["venkat_out"][:, 0]
    args = (x[:, 0], x[:, 1], x[:, 2], x[:, 3])
    assert parity.max_ulp(ma.probe_venkat(*args, arith=ma.ARITH_STRICT), ref) == 0
    fast = ma.probe_venkat(*args, arith=ma.ARITH_FAST)   # N / D with the common factor du cancelled
    assert float(np.abs(fast - ref).max()) < 1e-14


def test_vanalbada_limiter(lib, g):
    import miniaero_b200 as ma
    x, ref = g["vanalbada_in"], g["vanalbada_out"][:, 0]
    for arith in (ma.ARITH_STRICT, ma.ARITH_FAST):
        out = ma.probe_vanalbada(x[:, 0], x[:, 1], x[:, 2], arith=arith)
        if arith == ma.ARITH_STRICT:
            assert parity.max_ulp(out, ref) == 0
        else:
            assert float(np.abs(out - ref).max()) < 1e-15
