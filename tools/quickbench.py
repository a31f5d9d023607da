import sys, time, json
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miniaero_b200 as ma
def run(nx,ny,nz,second,visc,ptype=0,steps=5,tile=(0,0,0),bt=0,lx=0.3048,ly=1.0,lz=1.0,dt=5e-7):
    opt=ma.Options(problem_type=ptype,lx=lx,ly=ly,lz=lz,angle=0.0,nx=nx,ny=ny,nz=nz,ntimesteps=steps,dt=dt,second_order_space=second,viscous=visc)
    t=time.time(); mesh=ma.Parallel3DMesh.from_options(opt).fillMeshData(); tm=time.time()-t
    t=time.time(); s=ma.TimeSolverExplicitRK4(mesh,opt,tile_dims=tile,block_threads=bt); tl=time.time()-t
    s.initialize(); s.step(2); s.reset_timing(); s.step(steps)
    T=s.timing(); cu=T['cell_updates']/T['step_seconds']
    s.set_profiling(True); s.reset_timing(); s.step(2); P=s.timing()
    print(json.dumps(dict(n=(nx,ny,nz),second=second,visc=visc,tile=tile,bt=bt,mesh_s=round(tm,2),layout_s=round(tl,2),ms_per_step=round(1e3*T['step_seconds']/steps,3),cell_updates_per_s='%.3e'%cu,
        frac_roofline=round(cu*(4648 if second else 1928)/6549.4e9,4),grad_ms=round(1e3*P['grad_seconds']/2,3),flux_ms=round(1e3*P['flux_seconds']/2,3),dev_GB=round(T['device_bytes']/1e9,2),tiles=T['num_tiles'])),flush=True)
if __name__=='__main__':
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == 'variants':
        import os
        for gv, fv, tile, bt in [('gather','gather',(8,8,4),0), ('gather','gather',(8,8,8),0), ('gather','gather',(8,4,4),0), ('gather','gather',(8,8,4),128),
                                 ('tile','gather',(8,8,4),0), ('tile','gather',(8,4,4),0), ('tile','gather',(8,8,8),0),
                                 ('gather','tile',(8,4,4),128), ('tile','tile',(8,4,4),128)]:
            os.environ['MINIAERO_GRAD_KERNEL'] = gv; os.environ['MINIAERO_FLUX_KERNEL'] = fv
            print(gv, fv, end=' ')
            try:
                run(256,256,128,1,1,tile=tile,bt=bt)
            except Exception as e:
                print('FAILED', tile, bt, e, flush=True)
    elif len(sys.argv) > 1 and sys.argv[1] == 'tiles':
        for tile, bt in [((0,0,0),0), ((8,4,4),256), ((4,4,8),128), ((8,8,4),256), ((8,8,4),128), ((16,4,4),256), ((8,8,8),256), ((4,4,4),128), ((4,4,4),64)]:
            try:
                run(256,256,128,1,1,tile=tile,bt=bt)
            except Exception as e:
                print('FAILED', tile, bt, e, flush=True)
        run(256,256,128,1,0)
        run(256,256,128,0,0)
        run(256,128,64,1,1,ptype=1,lx=2.0,ly=0.008,lz=1.0,dt=3e-8)
    else:
        run(128,128,128,1,1)
        run(128,128,128,0,0)
        run(256,256,128,1,1)
        run(256,256,128,1,0)
