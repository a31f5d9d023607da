"""Parity norms (SURVEY.md §7 hard part 2).

The tolerance of the north star — 1e-12 relative per step, 1e-10 relative L2/Linf after 100 steps — is
applied on FIELD scales: rho and rho*E are normalised by their own Linf / L2 norm, the three momentum
components by the norm of the momentum VECTOR field, because components that are physically zero (rho*v,
rho*w in Sod; rho*w on the flat plate) hold only roundoff and have no scale of their own.
"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TOL_PER_STEP = 1e-12
TOL_100_STEPS = 1e-10


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def field_errors(test, ref):
    """Returns (linf, l2): worst relative error over the 5 conserved fields, field-scale normalised."""
    test = np.asarray(test, dtype=np.float64).reshape(-1, 5)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1, 5)
    assert test.shape == ref.shape
    d = test - ref
    mom = np.sqrt((ref[:, 1:4] ** 2).sum(axis=1))
    linf_scale = np.array([np.abs(ref[:, 0]).max(), mom.max(), mom.max(), mom.max(), np.abs(ref[:, 4]).max()])
    l2_scale = np.array([np.linalg.norm(ref[:, 0]), np.linalg.norm(mom), np.linalg.norm(mom), np.linalg.norm(mom),
                         np.linalg.norm(ref[:, 4])])
    # a field that is identically zero (momentum at t=0) has no scale: fall back to the energy-free unit
    linf_scale = np.where(linf_scale > 0, linf_scale, 1.0)
    l2_scale = np.where(l2_scale > 0, l2_scale, 1.0)
    linf = (np.abs(d).max(axis=0) / linf_scale).max()
    l2 = (np.linalg.norm(d, axis=0) / l2_scale).max()
    return float(linf), float(l2)


def field_error_parts(test, ref):
    """The pieces of field_errors for ONE block of a partitioned field: per-field max |error|, sum of squared
    errors, Linf scale and squared L2 scale (rho, momentum vector x3, rho*E).  combine_parts() turns the blocks'
    pieces into the norms of the global field."""
    test = np.asarray(test, dtype=np.float64).reshape(-1, 5)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1, 5)
    assert test.shape == ref.shape
    d = test - ref
    mom2 = (ref[:, 1:4] ** 2).sum(axis=1)
    linf_scale = [float(np.abs(ref[:, 0]).max()), float(np.sqrt(mom2.max())), float(np.abs(ref[:, 4]).max())]
    l2_scale2 = [float((ref[:, 0] ** 2).sum()), float(mom2.sum()), float((ref[:, 4] ** 2).sum())]
    return {"err_max": np.abs(d).max(axis=0).tolist(), "err_sq": (d ** 2).sum(axis=0).tolist(),
            "linf_scale": linf_scale, "l2_scale2": l2_scale2}


def combine_parts(parts):
    """(linf, l2) of the global field from the per-block pieces: a block is part of ONE field, so its errors are
    measured on that field's scale — a block the wave has not reached yet (momentum ~1e-7 of the global maximum)
    has no momentum scale of its own, exactly like the zero components of field_errors."""
    group = [0, 1, 1, 1, 2]
    linf_scale = [max(p["linf_scale"][g] for p in parts) for g in range(3)]
    l2_scale = [np.sqrt(sum(p["l2_scale2"][g] for p in parts)) for g in range(3)]
    linf_scale = [s if s > 0 else 1.0 for s in linf_scale]
    l2_scale = [s if s > 0 else 1.0 for s in l2_scale]
    linf = max(max(p["err_max"][k] for p in parts) / linf_scale[group[k]] for k in range(5))
    l2 = max(np.sqrt(sum(p["err_sq"][k] for p in parts)) / l2_scale[group[k]] for k in range(5))
    return float(linf), float(l2)


def max_ulp(test, ref):
    """Largest distance in units in the last place between two float64 arrays (0 = bit identical,
    treating +0 and -0 as equal)."""
    a = np.ascontiguousarray(test, dtype=np.float64).ravel()
    b = np.ascontiguousarray(ref, dtype=np.float64).ravel()
    ia = a.view(np.int64).copy()
    ib = b.view(np.int64).copy()
    ia = np.where(ia < 0, np.int64(-2**63) - ia, ia)
    ib = np.where(ib < 0, np.int64(-2**63) - ib, ib)
    return int(np.abs(ia - ib).max()) if a.size else 0
