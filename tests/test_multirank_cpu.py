"""N > 1 host logic on CPU: world_size-2 and -4 runs over the gloo backend (no GPU)."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,dims,ptype", [(2, (12, 6, 4), 0), (2, (4, 10, 6), 1), (4, (8, 8, 3), 2), (8, (8, 6, 4), 0)])
def test_block_decomposition_and_ghost_lists(lib, world, dims, ptype):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "_multirank_worker.py")] + [str(d) for d in dims] + [str(ptype)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "MULTIRANK_OK world=%d" % world in p.stdout


def test_processor_arrangement_matches_the_reference_rule(lib):
    """compute_processor_arrangement (Parallel3DMesh.C:247-303): bisect the largest remaining dimension
    (ties -> x, then y, then z); rank = x + npx*(y + npy*z)."""
    import miniaero_b200 as ma

    def ref_rule(n, nranks):
        rem, npd = list(n), [1, 1, 1]
        k = nranks
        while k > 1:
            d = max(range(3), key=lambda i: (rem[i], -i))
            rem[d] = rem[d] / 2.0
            npd[d] *= 2
            k //= 2
        return tuple(npd)

    for n, r in [((1024, 512, 256), 2), ((1024, 1024, 256), 4), ((1024, 1024, 512), 8), ((16, 16, 16), 8),
                 ((128, 4, 4), 4), ((64, 32, 2), 16)]:
        opt = ma.Options(nx=n[0], ny=n[1], nz=n[2])
        m = ma.Parallel3DMesh.from_options(opt, r - 1, r)
        h_nproc = None
        # only the decomposition is needed: generate a small stand-in when the mesh is large
        if n[0] * n[1] * n[2] > 1 << 16:
            s = max(n) // 16
            opt = ma.Options(nx=n[0] // s, ny=n[1] // s, nz=max(1, n[2] // s))
            m = ma.Parallel3DMesh.from_options(opt, r - 1, r)
            n = (opt.nx, opt.ny, opt.nz)
        md = m.fillMeshData()
        assert tuple(md.nproc) == ref_rule(n, r), (n, r, md.nproc)
        bx, by, bz = md.block
        assert r - 1 == bx + md.nproc[0] * (by + md.nproc[1] * bz)


def test_block_decomposition_entry_point_matches_the_mesh(lib):
    """ma_block_decomposition (what the structured constructor and the host driver use) == the decomposition
    ma_mesh_generate reports, for every rank of several arrangements; and the reference's error for non-2^k ranks."""
    import ctypes as C
    import miniaero_b200 as ma
    from miniaero_b200 import _abi
    for n, r in [((16, 8, 8), 1), ((16, 8, 8), 2), ((12, 12, 6), 4), ((8, 8, 8), 8), ((32, 4, 2), 4)]:
        opt = ma.Options(nx=n[0], ny=n[1], nz=n[2])
        for rank in range(r):
            md = ma.Parallel3DMesh.from_options(opt, rank, r).fillMeshData()
            a = [(C.c_int * 3)() for _ in range(4)]
            _abi.check(lib.ma_block_decomposition(C.byref(opt), rank, r, *a))
            assert [tuple(x) for x in a] == [tuple(md.nproc), tuple(md.block), tuple(md.nlocal), tuple(md.offset)]
    opt = ma.Options(nx=8, ny=8, nz=8)
    a = [(C.c_int * 3)() for _ in range(4)]
    assert lib.ma_block_decomposition(C.byref(opt), 0, 3, *a) != 0
    assert b"power of 2" in lib.ma_last_error()
