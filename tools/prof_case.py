"""Small driver for ncu captures: one workload, a few RK4 steps through the C ABI (no timing claims).
    python tools/prof_case.py NX NY NZ SECOND VISCOUS [PROBLEM_TYPE] [STEPS] [ARITH]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miniaero_b200 as ma  # noqa: E402


def main():
    a = sys.argv[1:]
    nx, ny, nz, second, visc = (int(x) for x in a[:5])
    ptype = int(a[5]) if len(a) > 5 else 0
    steps = int(a[6]) if len(a) > 6 else 3
    arith = int(a[7]) if len(a) > 7 else ma.ARITH_FAST
    geo = {0: (0.3048, 1.0, 1.0, 5e-7), 1: (2.0, 0.032, 1.0, 3e-8), 2: (2.0, 2.0, 1.0, 1e-6)}[ptype]
    opt = ma.Options(problem_type=ptype, lx=geo[0], ly=geo[1], lz=geo[2], angle=30.0 if ptype == 2 else 0.0, nx=nx,
                     ny=ny, nz=nz, ntimesteps=steps, dt=geo[3], second_order_space=second, viscous=visc)
    mesh = ma.Parallel3DMesh.from_options(opt).fillMeshData()
    tile = tuple(int(x) for x in os.environ.get('MINIAERO_TILE', '0,0,0').split(','))
    s = ma.TimeSolverExplicitRK4(mesh, opt, arith=arith, tile_dims=tile)
    s.initialize()
    s.step(steps)
    t = s.timing()
    print("cells %d steps %d ms/step %.3f launches %d" % (nx * ny * nz, steps, 1e3 * t["step_seconds"] / steps,
                                                         t["kernel_launches"]))


if __name__ == "__main__":
    main()
