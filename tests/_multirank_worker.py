"""Worker of tests/test_multirank_cpu.py (launched under torch.distributed.run, gloo backend, CPU only):
checks the host side of the N > 1 path — block decomposition, ghost lists and their pairing across ranks —
with the same exchange pattern the NCCL halo uses (ascending peer rank, one send + one recv per peer)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniaero_b200 as ma  # noqa: E402


def exchange(rank, mesh, payload_of):
    """Emulates one halo round: for each peer (ascending rank) send payload[send ids] / recv into ghosts."""
    out = {}
    so = ro = 0
    reqs, bufs = [], []
    for p in range(mesh.num_ranks):
        sc, rc = int(mesh.sendCount[p]), int(mesh.recvCount[p])
        if p != rank and sc:
            t = torch.from_numpy(np.ascontiguousarray(payload_of(mesh.send_local_ids[so:so + sc])))
            reqs.append(dist.isend(t, p))
            bufs.append(t)
        if p != rank and rc:
            shape = payload_of(mesh.recv_local_ids[ro:ro + rc]).shape
            t = torch.empty(shape, dtype=torch.float64)
            reqs.append(dist.irecv(t, p))
            out[p] = (mesh.recv_local_ids[ro:ro + rc].copy(), t)
        so += sc
        ro += rc
    for r in reqs:
        r.wait()
    return out


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    nx, ny, nz, ptype = (int(x) for x in sys.argv[1:5])
    opt = ma.Options(problem_type=ptype, lx=2.0, ly=1.0, lz=0.5, angle=15.0 if ptype == 2 else 0.0, nx=nx, ny=ny, nz=nz)
    mesh = ma.Parallel3DMesh.from_options(opt, rank, world).fillMeshData()
    n_own, n_gh = mesh.num_owned_cells, mesh.num_ghosts
    assert mesh.num_ranks == world and mesh.my_rank == rank
    assert np.prod(mesh.nproc) == world and np.prod(mesh.nlocal) == n_own
    gids = mesh.global_ids.astype(np.int64)

    # 1. owned cells partition the global mesh
    counts = [None] * world
    dist.all_gather_object(counts, gids[:n_own].tolist())
    allg = np.concatenate([np.asarray(c) for c in counts])
    assert len(allg) == nx * ny * nz and len(np.unique(allg)) == nx * ny * nz

    # 2. ghosts: the ids a peer sends are exactly the ghosts this rank expects from it, in the same order
    assert int(mesh.recvCount.sum()) == n_gh and (mesh.recv_local_ids >= n_own).all()
    assert (mesh.send_local_ids < n_own).all()
    got = exchange(rank, mesh, lambda ids: gids[ids].astype(np.float64))
    seen = 0
    for p, (ids, t) in got.items():
        assert np.array_equal(t.numpy().astype(np.int64), gids[ids]), "ghost pairing with rank %d" % p
        seen += len(ids)
    assert seen == n_gh

    # 3. ghost geometry equals the owner's, bit for bit (coordinates and volumes)
    geo = np.concatenate([mesh.cell_coordinates, mesh.cell_volumes[:, None]], axis=1)
    got = exchange(rank, mesh, lambda ids: geo[ids])
    for p, (ids, t) in got.items():
        assert np.array_equal(t.numpy().view(np.int64), geo[ids].view(np.int64)), "ghost geometry from rank %d" % p

    # 4. a face between an owned cell and a ghost has the owned cell as elem1 (SURVEY §7 hard part 7) and each
    #    ghost touches exactly one owned cell (face neighbours only, Parallel3DMesh.C:98-174); faces between two
    #    ghosts of one layer are kept, as the reference keeps them (MeshProcessor.C:173-186 only drops the
    #    neighbour-less faces of ghosts) — no owned cell ever reads them
    conn = mesh.internal_faces.face_cell_conn_
    gg = (conn[:, 0] >= n_own) & (conn[:, 1] >= n_own)
    gh = (conn[:, 1] >= n_own) & ~gg
    assert (conn[~gg, 0] < n_own).all()
    assert len(np.unique(conn[gh, 1])) == n_gh == int(gh.sum())

    # 5. faces of the whole decomposition add up to the single-domain mesh
    nint_local = int((~gh & ~gg).sum()) + 0.5 * int(gh.sum())
    nb_local = sum(f.nfaces_ for _, f in mesh.boundary_faces)
    tot = torch.tensor([nint_local, nb_local], dtype=torch.float64)
    dist.all_reduce(tot)
    assert tot[0].item() == (nx - 1) * ny * nz + nx * (ny - 1) * nz + nx * ny * (nz - 1)
    assert tot[1].item() == 2 * (nx * ny + ny * nz + nx * nz)
    dist.barrier()
    if rank == 0:
        print("MULTIRANK_OK world=%d blocks=%s" % (world, mesh.nproc))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
