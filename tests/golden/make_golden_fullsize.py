"""Generates tests/golden/sod_fullsize_column.npz: the reference (oracle/_ref/miniAero.cell) on a 512 x 4 x 4 mesh with
the cell size, time step and physics of the 512 x 512 x 256 benchmark mesh (BASELINE configs[1], bench.py's `sod_o2`
workload: second order, inviscid).  The 3-D Sod problem is one-dimensional: every x-line of the 67 M-cell run must
reproduce this line — the size-independent property tests/test_gpu_fullsize.py checks at the benchmark's full size.
(With the viscous term on the property does not hold in the reference itself: its slip walls carry no viscous stress,
so the wall-adjacent lines pick up a transverse momentum of 1e-6 relative.)

Run here (needs a built oracle/_ref):  python tests/golden/make_golden_fullsize.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refrun  # noqa: E402

NX, FULL = 512, (512, 512, 256)
INP = dict(problem_type=0, lx=0.3048, ly=1.0 * 4 / FULL[1], lz=1.0 * 4 / FULL[2], angle=0.0, nx=NX, ny=4, nz=4,
           dt=5e-7, output_results=1, output_frequency=100000, second_order=1, viscous=0)


def main():
    out = {}
    for n in (1, 2):
        o = refrun.run_reference(dict(INP, ntimesteps=n), kind="cell")
        sol = refrun.solution_from_dumps(o["dumps"]).reshape(NX, 4, 4, 5)
        line = sol[:, 1, 1, :]
        spread = np.abs(sol - line[:, None, None, :]).max(axis=(1, 2))
        out["line_step%d" % n] = line
        out["spread_step%d" % n] = spread   # the reference's own deviation between its 16 x-lines (roundoff)
        print(n, "steps: max spread between the reference's x-lines", spread.max(axis=0))
    np.savez_compressed(os.path.join(HERE, "sod_fullsize_column.npz"), **out)


if __name__ == "__main__":
    main()
