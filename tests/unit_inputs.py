"""Seeded inputs of the device-function parity tests (tests/golden/unit_functions.npz holds them together with the
reference's outputs, made by tests/golden/make_golden_unit.py from oracle/_ref/unit_oracle)."""
import numpy as np

N = 1024
RGAS, GAMMA = 287.05, 1.4


def _states(rng, n):
    """plausible primitive states (rho, u, v, w, T)"""
    rho = rng.uniform(0.05, 2.0, n)
    vel = rng.uniform(-600.0, 600.0, (n, 3))
    T = rng.uniform(150.0, 900.0, n)
    return np.column_stack([rho, vel, T])


def _frames(rng, n):
    """area-weighted normal a, unit tangent t, binormal a x t (Face.C:81-96's orthogonal frame)"""
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    helper = rng.normal(size=(n, 3))
    t = np.cross(nrm, helper)
    t /= np.linalg.norm(t, axis=1)[:, None]
    area = 10.0 ** rng.uniform(-6.0, -2.0, n)
    a = nrm * area[:, None]
    return a, t, np.cross(a, t)


def make_inputs():
    rng = np.random.default_rng(20261017)
    out = {}
    # ---- Roe flux: generic pairs, identical states (Extrapolate_BC), small jumps, and states whose normal Mach
    # number sits near 0 and near +-1 so that every branch of the eigenvalue fix (Roe_Flux.h:164-178) is taken
    vl, vr = _states(rng, N), _states(rng, N)
    a, t, b = _frames(rng, N)
    q = N // 4
    vr[:q] = vl[:q]
    vr[q:2 * q] = vl[q:2 * q] * (1.0 + 1e-3 * rng.normal(size=(q, 5)))
    unit = a / np.linalg.norm(a, axis=1)[:, None]
    c = np.sqrt(GAMMA * RGAS * vl[:, 4])
    for lo, mach in ((2 * q, 0.0), (2 * q + q // 3, 1.0), (2 * q + 2 * (q // 3), -1.0)):
        sl = slice(lo, lo + q // 3)
        m = mach + rng.uniform(-0.12, 0.12, q // 3)
        tang = np.cross(unit[sl], rng.normal(size=(q // 3, 3))) * 30.0
        vl[sl, 1:4] = unit[sl] * (m * c[sl])[:, None] + tang
        vr[sl] = vl[sl] * (1.0 + 1e-2 * rng.normal(size=(q // 3, 5)))
    out["roe"] = np.column_stack([vl, vr, a, t, b])
    # ---- viscous flux
    g = rng.normal(size=(N, 15)) * 10.0 ** rng.uniform(0.0, 5.0, (N, 1))
    a, _, _ = _frames(rng, N)
    out["viscous"] = np.column_stack([g, _states(rng, N), a])
    # ---- primitives from conservative states
    v = _states(rng, N)
    e = RGAS / (GAMMA - 1.0) * v[:, 4] + 0.5 * (v[:, 1:4] ** 2).sum(axis=1)
    out["primitives"] = np.column_stack([v[:, 0], v[:, 0:1] * v[:, 1:4], v[:, 0] * e])
    # ---- limiters: dumax >= 0 >= dumin; du of both signs, exactly zero, below the 1e-40 / DBL_EPSILON thresholds
    dumax = np.abs(rng.normal(size=N)) * 10.0 ** rng.uniform(-8.0, 3.0, N)
    dumin = -np.abs(rng.normal(size=N)) * 10.0 ** rng.uniform(-8.0, 3.0, N)
    du = rng.normal(size=N) * 10.0 ** rng.uniform(-10.0, 3.0, N)
    du[:64] = 0.0
    du[64:128] = rng.choice([-1.0, 1.0], 64) * 10.0 ** rng.uniform(-45.0, -38.0, 64)
    du[128:192] = rng.choice([-1.0, 1.0], 64) * 10.0 ** rng.uniform(-17.0, -15.0, 64)
    dumax[192:256] = 0.0
    dumin[256:320] = 0.0
    dx2 = 10.0 ** rng.uniform(-12.0, -2.0, N)
    out["venkat"] = np.column_stack([dumax, dumin, du, dx2])
    out["vanalbada"] = np.column_stack([dumax, dumin, du])
    return out
