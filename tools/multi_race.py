"""Race hunt for the block-decomposed step (launch under torch.distributed.run): repeats a short multi-rank run under
kernel / overlap variants and reports, per variant, the worst per-rank error against the reference's MPI build."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import cases  # noqa: E402
import parity  # noqa: E402
import miniaero_b200 as ma  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    comm = ma.HaloComm.from_torch_distributed(local)
    name = "sod_o2_visc"
    inp = cases.PARALLEL[name][0]
    g = np.load(os.path.join(parity.GOLDEN, "par_%s_%d.npz" % (name, world)))
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    out = {}
    detail = None
    for tile, arith in [((0, 0, 0), 0), ((8, 8, 8), 0), ((16, 4, 8), 0), ((4, 4, 8), 1), ((4, 4, 4), 0), ((2, 2, 2), 0)]:
        for n in (2,):
            opt = ma.Options(**cases.opts_kwargs(dict(inp, ntimesteps=n)))
            mesh = ma.Parallel3DMesh.from_options(opt, rank, world).fillMeshData()
            solver = ma.TimeSolverExplicitRK4(mesh, opt, device=local, arith=arith, comm=comm, tile_dims=tile)
            solver.initialize()
            solver.step(n)
            sol = solver.solution()
            ref = g["r%d_step%d" % (rank, n)]
            linf, _ = parity.field_errors(sol, ref)
            rows = [None] * world
            dist.all_gather_object(rows, (linf, 0))
            out["tile%s/arith%d/step%d" % (tile, arith, n)] = rows
            if tile == (0, 0, 0) and rank == 2:
                err = np.abs(sol - ref)
                idx = np.argsort(err.max(axis=1))[::-1][:6]
                gids = mesh.global_ids[:mesh.num_owned_cells]
                detail = ["cell %d gid %d xyz %s err %s sol %s ref %s" % (i, gids[i], np.array2string(np.asarray(mesh.cell_coordinates[i]), precision=4), np.array2string(err[i], precision=2), np.array2string(sol[i], precision=8), np.array2string(ref[i], precision=8)) for i in idx]
                detail.append("scale: max|mom| %.3e max rho %.3e nbad cells %d of %d" % (np.abs(ref[:, 1:4]).max(), ref[:, 0].max(), int((err.max(axis=1) > 1e-10 * np.abs(ref).max(axis=1)).sum()), len(ref)))
            del solver
    allrows = [None] * world
    dist.all_gather_object(allrows, detail)
    if rank == 0:
        for d in allrows:
            if d:
                print("\n".join(d))
    if rank == 0:
        for k, v in out.items():
            print(k, " ".join("%.1e(%d)" % r for r in v), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
