"""Worker of tests/test_gpu_multi.py: one process per GPU (torch.distributed.run, NCCL), the block-decomposed
RK4 step with NCCL halo exchange through the C ABI, checked per rank against the full-precision results of the
reference's WITH_MPI build (tests/golden/par_*.npz, made by tests/golden/make_golden.py parallel)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import cases  # noqa: E402
import parity  # noqa: E402
import miniaero_b200 as ma  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    comm = ma.HaloComm.from_torch_distributed(local)
    report = {}
    for name in sys.argv[1:]:
        inp, rank_counts, _, rel_tol, floor = cases.PARALLEL[name]
        if world not in rank_counts:
            continue
        g = np.load(os.path.join(parity.GOLDEN, "par_%s_%d.npz" % (name, world)))
        serial = parity.golden(name) if os.path.isfile(os.path.join(parity.GOLDEN, name + ".npz")) else None
        for arith, overlap in ((ma.ARITH_STRICT, True), (ma.ARITH_STRICT, False), (ma.ARITH_FAST, True)):
            for n in (1, 2, 100):
                opt = ma.Options(**cases.opts_kwargs(dict(inp, ntimesteps=n)))
                mesh = ma.Parallel3DMesh.from_options(opt, rank, world).fillMeshData()
                solver = ma.TimeSolverExplicitRK4(mesh, opt, device=local, arith=arith, comm=comm, overlap_halo=overlap)
                solver.initialize()
                solver.step(n)
                sol = solver.solution()
                ref = g["r%d_step%d" % (rank, n)]
                # the blocks are pieces of one global field: errors are measured on the global field's scale
                entry = {"ulp": parity.max_ulp(sol, ref), "parts": parity.field_error_parts(sol, ref),
                         "launches": solver.timing()["kernel_launches"]}
                if serial is not None and ("cell_step%d" % n) in serial:
                    gids = mesh.global_ids[:mesh.num_owned_cells]
                    entry["parts_vs_single_domain"] = parity.field_error_parts(sol, serial["cell_step%d" % n][gids])
                rows = [None] * world
                dist.all_gather_object(rows, entry)
                linf, l2 = parity.combine_parts([r["parts"] for r in rows])
                summary = {"ulp": max(r["ulp"] for r in rows), "linf": linf, "l2": l2,
                           "launches": [r["launches"] for r in rows]}
                if "parts_vs_single_domain" in entry:
                    summary["linf_vs_single_domain"] = parity.combine_parts([r["parts_vs_single_domain"] for r in rows])[0]
                report["%s/arith%d/overlap%d/step%d" % (name, arith, int(overlap), n)] = [summary]
                del solver
        # the structured constructor (block layout from (i, j, k), geometry evaluated on the device): same bits
        opt = ma.Options(**cases.opts_kwargs(dict(inp, ntimesteps=2)))
        solver = ma.TimeSolverExplicitRK4.from_options(opt, rank, world, device=local, arith=ma.ARITH_STRICT, comm=comm)
        solver.initialize()
        solver.step(2)
        ulp = parity.max_ulp(solver.solution(), g["r%d_step2" % rank])
        rows = [None] * world
        dist.all_gather_object(rows, ulp)
        report["%s/structured/arith1/step2" % name] = [{"ulp": max(rows)}]
        del solver
        # the reference's own parallel integration test: results.<rank> vs results.<rank>.gold
        if ("r%d_gold" % rank) in g:
            import refrun
            opt = ma.Options(**cases.opts_kwargs(inp))
            mesh = ma.Parallel3DMesh.from_options(opt, rank, world).fillMeshData()
            solver = ma.TimeSolverExplicitRK4(mesh, opt, device=local, comm=comm)
            tmp = "/tmp/miniaero_results_%d.%d" % (os.getpid(), rank)
            solver.Solve_to(tmp)
            bad = refrun.numeric_text_diff(np.loadtxt(tmp), g["r%d_gold" % rank], rel_tol, floor)
            os.remove(tmp)
            rows = [None] * world
            dist.all_gather_object(rows, bad)
            report["%s/gold_diff_lines" % name] = rows
    if rank == 0:
        print("MULTIGPU_REPORT " + json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
