"""Parity at the benchmark's FULL size (BASELINE configs[1]: second-order Sod, 512 x 512 x 256 = 67 108 864 cells)
through a size-independent property: the 3-D Sod problem is one-dimensional, so every one of the 131 072 x-lines of
the big run must reproduce the reference's x-line for the same cell size and time step
(tests/golden/sod_fullsize_column.npz, made by tests/golden/make_golden_fullsize.py from the unmodified reference on
a 512 x 4 x 4 mesh; the reference's own 16 lines agree with each other to roundoff).  Tolerance: the north star's
1e-12 relative per step on field scales."""
import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu

FULL = (512, 512, 256)


def _enough_memory():
    try:
        import psutil
        import torch
        free, _ = torch.cuda.mem_get_info(0)
        return psutil.virtual_memory().available > 90e9 and free > 60e9
    except Exception:
        return False


def test_every_x_line_of_the_full_size_sod_run_is_the_reference_line(lib):
    if not _enough_memory():
        pytest.skip("needs ~90 GB of host memory and ~60 GB of device memory")
    import miniaero_b200 as ma
    g = parity.golden("sod_fullsize_column")
    nx, ny, nz = FULL
    opt = ma.Options(problem_type=0, lx=0.3048, ly=1.0, lz=1.0, angle=0.0, nx=nx, ny=ny, nz=nz, ntimesteps=2, dt=5e-7,
                     second_order_space=1, viscous=0)
    mesh = ma.Parallel3DMesh.from_options(opt).fillMeshData()
    solver = ma.TimeSolverExplicitRK4(mesh, opt)      # FAST arithmetic: the benchmarked path
    solver.release_mesh()
    del mesh
    solver.initialize()
    for n in (1, 2):
        solver.step(1)
        sol = solver.solution().reshape(nx, ny, nz, 5)   # cell id = (i * ny + j) * nz + k (Parallel3DMesh.h:462-464)
        line = g["line_step%d" % n]
        mom = np.sqrt((line[:, 1:4] ** 2).sum(axis=1)).max()
        scale = np.array([np.abs(line[:, 0]).max(), mom, mom, mom, np.abs(line[:, 4]).max()])
        worst = np.zeros(5)
        for i0 in range(0, nx, 32):   # slabs keep the temporaries small
            d = np.abs(sol[i0:i0 + 32] - line[i0:i0 + 32, None, None, :])
            worst = np.maximum(worst, d.max(axis=(0, 1, 2)))
        rel = worst / scale
        print("step", n, "worst relative deviation of any x-line from the reference line:", rel)
        assert (rel <= n * parity.TOL_PER_STEP).all(), rel
        assert np.isfinite(sol).all()
