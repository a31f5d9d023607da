"""Inputs of the parity cases.

`REFERENCE_TESTS` are the reference's own serial integration tests (tests/*/miniaero.inp)
with their gold-file tolerances; `EXTRA` widens the template-instantiation matrix of
compute_face_flux<second_order, viscous> (TimeSolverExplicitRK4.h:402-426) on small meshes.
"""

def _inp(problem_type, lx, ly, lz, angle, nx, ny, nz, ntimesteps, dt, second_order, viscous):
    return dict(problem_type=problem_type, lx=lx, ly=ly, lz=lz, angle=angle, nx=nx, ny=ny, nz=nz,
                ntimesteps=ntimesteps, dt=dt, output_results=1, output_frequency=100000,
                second_order=second_order, viscous=viscous)


# name -> (inp, gold file relative to the reference tests dir, rel_tol, floor)
REFERENCE_TESTS = {
    "3D_Sod_Serial": (_inp(0, 0.3048, 1.0, 1.0, 0.0, 128, 4, 4, 100, 2e-6, 0, 0), 1e-3, 1e-6),
    "Ramp_Serial": (_inp(2, 2.0, 2.0, 1.0, 30.0, 64, 32, 2, 400, 1e-5, 1, 0), 1e-2, 1e-3),
    "FlatPlate_Serial": (_inp(1, 2.0, 0.002, 1.0, 0.0, 16, 32, 2, 400, 3e-8, 1, 1), 1e-2, 1e-3),
}

# small cases covering every (second_order, viscous) combination and every boundary type
EXTRA = {
    "sod_o2": _inp(0, 0.3048, 1.0, 1.0, 0.0, 128, 4, 4, 100, 2e-6, 1, 0),        # tests/3D_Sod_Parallel's input
    "sod_o2_visc": _inp(0, 0.3048, 1.0, 1.0, 0.0, 64, 8, 8, 100, 2e-6, 1, 1),    # BASELINE config 4's physics
    "sod_o1_visc": _inp(0, 0.3048, 1.0, 1.0, 0.0, 64, 4, 4, 100, 2e-6, 0, 1),
    "flatplate_o1": _inp(1, 2.0, 0.002, 1.0, 0.0, 16, 32, 2, 100, 3e-8, 0, 0),   # NoSlip stays viscous (…RK4.h:457-460)
    "ramp_odd": _inp(2, 1.7, 0.9, 1.3, 17.0, 13, 7, 5, 100, 1e-5, 1, 1),         # ragged sizes: partial tiles
    # degenerate meshes: a single cell (six boundary faces, no internal face), a pair, a pencil, a one-cell-thick slab
    "one_cell": _inp(0, 0.3048, 1.0, 1.0, 0.0, 1, 1, 1, 100, 2e-6, 1, 1),
    "two_cells": _inp(0, 0.3048, 1.0, 1.0, 0.0, 2, 1, 1, 100, 2e-6, 1, 1),
    "pencil_z": _inp(1, 2.0, 0.002, 1.0, 0.0, 1, 1, 5, 100, 1e-8, 1, 1),
    "slab_o1": _inp(2, 2.0, 2.0, 1.0, 30.0, 3, 1, 2, 100, 1e-5, 0, 0),
}


def all_cases():
    out = {k: v[0] for k, v in REFERENCE_TESTS.items()}
    out.update(EXTRA)
    return out


def opts_kwargs(inp):
    """cases.py dict -> miniaero_b200.Options keyword arguments."""
    d = dict(inp)
    d["second_order_space"] = d.pop("second_order")
    return d


# multi-rank cases: name -> (inp, [rank counts], reference test dir holding results.<rank>.gold or None, rel_tol, floor)
PARALLEL = {
    "3D_Sod_Parallel": (_inp(0, 0.3048, 1.0, 1.0, 0.0, 128, 4, 4, 100, 2e-6, 1, 0), [2, 4], "3D_Sod_Parallel", 1e-3, 1e-6),
    "FlatPlate_Parallel": (_inp(1, 2.0, 0.002, 1.0, 0.0, 32, 64, 2, 400, 1e-8, 1, 1), [2, 8], "FlatPlate_Parallel", 1e-2, 1e-3),
    "sod_o2_visc": (EXTRA["sod_o2_visc"], [2, 4, 8], None, 0, 0),
    "ramp_o2_visc": (_inp(2, 2.0, 2.0, 1.0, 30.0, 32, 16, 8, 100, 1e-5, 1, 1), [2, 4], None, 0, 0),
    "sod_o1": (_inp(0, 0.3048, 1.0, 1.0, 0.0, 64, 8, 4, 100, 2e-6, 0, 0), [2], None, 0, 0),
}
