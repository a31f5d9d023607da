"""Quick single-GPU throughput probes (not the benchmark of record; see bench.py).
    python tools/quickbench.py [sweep|big]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miniaero_b200 as ma  # noqa: E402


def run(nx, ny, nz, second, visc, ptype=0, steps=5, tile=(0, 0, 0), bt=0, lx=0.3048, ly=1.0, lz=1.0, dt=5e-7, tag=""):
    opt = ma.Options(problem_type=ptype, lx=lx, ly=ly, lz=lz, angle=0.0, nx=nx, ny=ny, nz=nz, ntimesteps=steps, dt=dt,
                     second_order_space=second, viscous=visc)
    t = time.time()
    if os.environ.get('MINIAERO_MESH_PATH', 'structured') == 'structured':   # layout from (i,j,k), geometry on the device
        tm = 0.0
        s = ma.TimeSolverExplicitRK4.from_options(opt, tile_dims=tile, block_threads=bt)
    else:
        mesh = ma.Parallel3DMesh.from_options(opt).fillMeshData()
        tm = time.time() - t
        t = time.time()
        s = ma.TimeSolverExplicitRK4(mesh, opt, tile_dims=tile, block_threads=bt)
    tl = time.time() - t
    s.initialize()
    s.step(2)
    s.reset_timing()
    s.step(steps)
    T = s.timing()
    cu = T['cell_updates'] / T['step_seconds']
    s.set_profiling(True)
    s.reset_timing()
    s.step(2)
    P = s.timing()
    print(json.dumps(dict(tag=tag, n=(nx, ny, nz), second=second, visc=visc, tile=tile, bt=bt, mesh_s=round(tm, 2),
                          layout_s=round(tl, 2), ms_per_step=round(1e3 * T['step_seconds'] / steps, 3),
                          cell_updates_per_s='%.3e' % cu,
                          frac_roofline=round(cu * (4648 if second else 1928) / 6551.4e9, 4),
                          grad_ms=round(1e3 * P['grad_seconds'] / 8, 3), flux_ms=round(1e3 * P['flux_seconds'] / 8, 3),
                          dev_GB=round(T['device_bytes'] / 1e9, 2), tiles=T['num_tiles'])), flush=True)
    del s


def variant(gv, fv):
    os.environ['MINIAERO_GRAD_KERNEL'] = gv
    os.environ['MINIAERO_FLUX_KERNEL'] = fv


if __name__ == '__main__':
    mode = sys.argv[1] if len(sys.argv) > 1 else 'sweep'
    if mode == 'sweep':
        N = (256, 256, 128)
        for gv, fv, tile, bt in [('gather', 'gather', (8, 8, 4), 0), ('tma', 'tma', (8, 4, 4), 0), ('tma', 'tma', (8, 4, 4), 128),
                                 ('tma', 'tma', (8, 4, 4), 160), ('tma', 'tma', (8, 4, 4), 192), ('tma', 'tma', (8, 8, 4), 0),
                                 ('tma', 'tma', (4, 4, 4), 0), ('tma', 'tma', (4, 4, 4), 64), ('tma', 'tma', (4, 4, 8), 0),
                                 ('tma', 'gather', (8, 8, 4), 0)]:
            variant(gv, fv)
            try:
                run(*N, 1, 1, tile=tile, bt=bt, tag=gv + '/' + fv)
            except Exception as e:
                print('FAILED', gv, fv, tile, bt, e, flush=True)
        variant('tma', 'tma')
        run(*N, 1, 0, tag='o2 inviscid')
        run(*N, 0, 0, tag='o1 inviscid')
        run(256, 128, 64, 1, 1, ptype=1, lx=2.0, ly=0.008, lz=1.0, dt=3e-8, tag='flatplate')
    elif mode == 'one':  # python tools/quickbench.py one NX NY NZ [tag]
        nx, ny, nz = (int(x) for x in sys.argv[2:5])
        tile = tuple(int(x) for x in os.environ.get('MINIAERO_TILE', '0,0,0').split(','))
        run(nx, ny, nz, 1, 1, tile=tile, tag=sys.argv[5] if len(sys.argv) > 5 else '')
    elif mode == 'big':
        run(512, 512, 256, 1, 1, tag='sod_o2_visc 67M')
