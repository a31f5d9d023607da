"""miniaero_b200 — B200-native explicit-RK4 finite-volume step of miniAero behind a C ABI.

Python host mirror of the reference's interface for the hot path (same names, argument meaning and
error behaviour), over `include/miniaero_b200.h` via ctypes:

    reference (C++)                                      here
    ---------------                                      ----
    Options::read_options_file()        Options.h:73     Options.read_options_file(path="miniaero.inp")
    Parallel3DMesh(nx,ny,nz,lx,ly,lz,type,angle)         Parallel3DMesh(...).fillMeshData() -> MeshData
        + fillMeshData(MeshData&)       Parallel3DMesh.h:58,173
    TimeSolverExplicitRK4(mesh, opts).Solve()            TimeSolverExplicitRK4(mesh, opts).Solve()
                                        TimeSolverExplicitRK4.h:164-166

All compute runs in hand-written CUDA kernels for sm_100a inside libminiaero_b200.so; there is no CPU
fallback and importing the solver without the built library raises.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import (ARITH_FAST, ARITH_STRICT, BC_EXTRAPOLATE, BC_INFLOW, BC_NAMES, BC_NOSLIP, BC_TANGENT,
                   FIELD_GRADIENT, FIELD_LIMITER, FIELD_STAGE_PRIMITIVES, LIMITER_VENKAT, LIMITER_VANALBADA, MiniAeroError)

__all__ = ["Options", "Parallel3DMesh", "MeshData", "Faces", "TimeSolverExplicitRK4", "HaloComm", "MiniAeroError",
           "ARITH_FAST", "ARITH_STRICT", "probe_roe_flux", "probe_viscous_flux", "probe_primitives",
           "probe_venkat", "probe_vanalbada", "write_results"]

_BC_TYPE_NAMES = {v: k for k, v in BC_NAMES.items()}


class Options(_abi.Options):
    """`struct Options` (Options.h:47-102)."""

    def __init__(self, **kw):
        super().__init__()
        _abi.load().ma_options_default(C.byref(self))
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError("Options has no field %r" % k)
            setattr(self, k, v)

    def read_options_file(self, path="miniaero.inp"):
        """Options.h:73-101.  Unlike the reference (which warns and continues with garbage) a missing or
        short file raises."""
        _abi.check(_abi.load().ma_options_read(os.fsencode(path), C.byref(self)))
        return self


def _np(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return np.zeros(shape, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype).reshape(shape)


class Faces:
    """Host view of `struct Faces<Device>` (Faces.h:42-71)."""

    def __init__(self, cfaces, owner=None):
        n = cfaces.nfaces
        self.nfaces_ = n
        self.coordinates_ = _np(cfaces.coordinates, (n, 3), np.float64)
        self.face_normal_ = _np(cfaces.face_normal, (n, 3), np.float64)
        self.face_tangent_ = _np(cfaces.face_tangent, (n, 3), np.float64)
        self.face_binormal_ = _np(cfaces.face_binormal, (n, 3), np.float64)
        self.face_cell_conn_ = _np(cfaces.face_cell_conn, (n, 2), np.int32)
        self.cell_flux_index_ = _np(cfaces.cell_flux_index, (n, 2), np.int32)
        self._owner = owner


class MeshData:
    """Host view of `struct MeshData<Device>` (MeshData.h:43-58) backed by a ma_mesh_storage."""

    def __init__(self, handle):
        self._lib = _abi.load()
        self._handle = handle
        self.c_mesh = self._lib.ma_mesh_view(handle).contents
        m = self.c_mesh
        self.num_owned_cells = m.num_owned_cells
        self.num_ghosts = m.num_ghosts
        n = m.num_owned_cells + m.num_ghosts
        self.cell_coordinates = _np(m.cell_coordinates, (n, 3), np.float64)
        self.cell_volumes = _np(m.cell_volumes, (n,), np.float64)
        self.internal_faces = Faces(m.internal_faces, self)
        self.boundary_faces = [(_BC_TYPE_NAMES[m.boundary_type[b]], Faces(m.boundary_faces[b], self))
                               for b in range(m.num_boundary_sets)]
        self.num_ranks, self.my_rank = m.num_ranks, m.my_rank
        self.sendCount = _np(m.send_count, (m.num_ranks,), np.int32)
        self.recvCount = _np(m.recv_count, (m.num_ranks,), np.int32)
        self.send_local_ids = _np(m.send_local_ids, (int(self.sendCount.sum()),), np.int32)
        self.recv_local_ids = _np(m.recv_local_ids, (int(self.recvCount.sum()),), np.int32)
        self.global_ids = _np(self._lib.ma_mesh_global_ids(handle), (n,), np.int32)
        a = [(C.c_int * 3)() for _ in range(4)]
        self._lib.ma_mesh_decomposition(handle, *a)
        self.nproc, self.block, self.nlocal, self.offset = [tuple(x) for x in a]

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            self._lib.ma_mesh_free(h)


class Parallel3DMesh:
    """Parallel3DMesh (Parallel3DMesh.h:56-58): in-code structured hex block as unstructured arrays."""

    def __init__(self, nx, ny, nz, lx, ly, lz, problem_type, angle=0.0, rank=0, num_ranks=1):
        self.opt = Options(nx=nx, ny=ny, nz=nz, lx=lx, ly=ly, lz=lz, problem_type=problem_type, angle=angle)
        self.rank, self.num_ranks = rank, num_ranks

    @classmethod
    def from_options(cls, opt, rank=0, num_ranks=1):
        return cls(opt.nx, opt.ny, opt.nz, opt.lx, opt.ly, opt.lz, opt.problem_type, opt.angle, rank, num_ranks)

    def fillMeshData(self):
        """Parallel3DMesh::fillMeshData (Parallel3DMesh.h:173-449)."""
        lib = _abi.load()
        h = C.c_void_p()
        _abi.check(lib.ma_mesh_generate(C.byref(self.opt), self.rank, self.num_ranks, C.byref(h)))
        return MeshData(h)


class HaloComm:
    """NCCL communicator for the ghost exchange (replaces MPI_COMM_WORLD of CopyGhost.C:41-79).

    `HaloComm.from_torch_distributed(device)` bootstraps the NCCL id over an initialised
    torch.distributed process group (any backend): plumbing only."""

    def __init__(self, unique_id, num_ranks, rank, device):
        self._lib = _abi.load()
        h = C.c_void_p()
        _abi.check(self._lib.ma_comm_create(bytes(unique_id), num_ranks, rank, device, C.byref(h)))
        self._handle = h
        self.rank, self.num_ranks = rank, num_ranks

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(_abi.MA_COMM_ID_BYTES)
        _abi.check(_abi.load().ma_comm_get_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, device):
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        ids = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        return cls(ids[0], world, rank, device)

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            self._lib.ma_comm_destroy(h)


def _as_c_mesh(mesh):
    if isinstance(mesh, MeshData):
        return mesh.c_mesh, mesh
    if isinstance(mesh, _abi.Mesh):
        return mesh, None
    raise TypeError("mesh must be a MeshData or a ctypes Mesh")


class TimeSolverExplicitRK4:
    """TimeSolverExplicitRK4<Device> (TimeSolverExplicitRK4.h:160-177).

    solver = TimeSolverExplicitRK4(mesh_data, options); solver.Solve()
    The mesh is copied, renumbered into tiles and uploaded in the constructor; every array handed back
    is in the caller's cell order."""

    def __init__(self, input_mesh_data, options, device=0, arith=ARITH_FAST, tile_dims=(0, 0, 0), block_threads=0,
                 comm=None, overlap_halo=True, stream=None, limiter=0, share_cut_faces=0):
        self._lib = _abi.load()
        self.options = options
        cmesh, self._mesh_keepalive = _as_c_mesh(input_mesh_data)
        self._mesh = input_mesh_data
        self.num_owned_cells = cmesh.num_owned_cells
        cfg = _abi.SolverConfig()
        self._lib.ma_solver_config_default(C.byref(cfg))
        cfg.device, cfg.arith = device, arith
        cfg.tile_dims[0], cfg.tile_dims[1], cfg.tile_dims[2] = tile_dims
        cfg.block_threads = block_threads
        cfg.comm = comm._handle if comm is not None else None
        cfg.overlap_halo = 1 if overlap_halo else 0
        cfg.stream = stream
        cfg.limiter = limiter
        cfg.share_cut_faces = share_cut_faces
        self._comm = comm
        h = C.c_void_p()
        _abi.check(self._lib.ma_solver_create(C.byref(cmesh), C.byref(options), C.byref(cfg), C.byref(h)))
        self._handle = h

    @classmethod
    def from_options(cls, options, rank=0, nranks=1, device=0, arith=ARITH_FAST, tile_dims=(0, 0, 0), block_threads=0,
                     comm=None, overlap_halo=True, stream=None, limiter=0, share_cut_faces=0):
        """Parallel3DMesh + fillMeshData + the constructor in one call (ma_solver_create_structured): the block's
        device layout is built straight from (i, j, k) and its geometry is evaluated on the GPU; the same solver,
        bit for bit, as TimeSolverExplicitRK4(Parallel3DMesh.from_options(options, rank, nranks).fillMeshData(), ...).
        There is no host mesh afterwards: Solve() cannot write `results.<rank>` (use the mesh constructor for that)."""
        self = cls.__new__(cls)
        self._lib = _abi.load()
        self.options = options
        self._mesh = self._mesh_keepalive = None
        cfg = _abi.SolverConfig()
        self._lib.ma_solver_config_default(C.byref(cfg))
        cfg.device, cfg.arith = device, arith
        cfg.tile_dims[0], cfg.tile_dims[1], cfg.tile_dims[2] = tile_dims
        cfg.block_threads = block_threads
        cfg.comm = comm._handle if comm is not None else None
        cfg.overlap_halo = 1 if overlap_halo else 0
        cfg.stream = stream
        cfg.limiter = limiter
        cfg.share_cut_faces = share_cut_faces
        self._comm = comm
        h = C.c_void_p()
        _abi.check(self._lib.ma_solver_create_structured(C.byref(options), rank, nranks, C.byref(cfg), C.byref(h)))
        self._handle = h
        owned, ghosts = C.c_int(), C.c_int()
        _abi.check(self._lib.ma_solver_num_cells(h, C.byref(owned), C.byref(ghosts)))
        self.num_owned_cells, self.num_ghosts = owned.value, ghosts.value
        return self

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            self._lib.ma_solver_destroy(h)

    def release_mesh(self):
        """Drop the host mesh (the constructor already copied what the device needs); afterwards Solve()
        cannot write `results.<rank>` (it needs the cell coordinates)."""
        self._mesh = None
        self._mesh_keepalive = None

    # -- the reference's one public method
    def Solve(self):
        """TimeSolverExplicitRK4::Solve (TimeSolverExplicitRK4.h:207-539): initial conditions, ntimesteps RK4
        steps, `results.<rank>` in the working directory when options.output_results is set.
        Returns the solution [num_owned_cells][5]."""
        rank = self._comm.rank if self._comm is not None else 0
        return self.Solve_to("results.%d" % rank if self.options.output_results else None)

    def Solve_to(self, results_path):
        _abi.check(self._lib.ma_solver_solve(self._handle))
        sol = self.solution()
        if results_path:
            write_results(results_path, self._mesh, sol)
        return sol

    # -- finer-grained control used by the benchmark and the parity tests
    def initialize(self):
        _abi.check(self._lib.ma_solver_initialize(self._handle))

    def step(self, nsteps=1):
        _abi.check(self._lib.ma_solver_step(self._handle, nsteps))

    def synchronize(self):
        _abi.check(self._lib.ma_solver_synchronize(self._handle))
        self._inflight = []   # the copies of every submitted member are complete: their buffers may go

    def solution(self, out=None):
        if out is None:
            out = np.empty((self.num_owned_cells, 5), dtype=np.float64)
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == self.num_owned_cells * 5
        _abi.check(self._lib.ma_solver_get_solution(self._handle, out.ctypes.data))
        return out

    def solution_into(self, ptr):
        """Download into a raw host pointer (e.g. pinned memory owned by the caller)."""
        _abi.check(self._lib.ma_solver_get_solution(self._handle, ptr))

    def set_solution(self, U):
        if isinstance(U, int):
            _abi.check(self._lib.ma_solver_set_solution(self._handle, U))
            return
        U = np.ascontiguousarray(U, dtype=np.float64)
        assert U.size == self.num_owned_cells * 5
        _abi.check(self._lib.ma_solver_set_solution(self._handle, U.ctypes.data))

    def submit(self, state_in, state_out, nsteps=1):
        """Queue one ensemble member: upload `state_in`, advance `nsteps` RK4 steps, download into `state_out`
        (numpy arrays or raw host pointers; pinned memory lets the copies overlap the stepping of the neighbouring
        members).  Returns immediately; call synchronize() before reading `state_out`."""
        def ptr(a):
            if isinstance(a, int):
                return a
            assert a.dtype == np.float64 and a.flags.c_contiguous and a.size == self.num_owned_cells * 5
            return a.ctypes.data
        _abi.check(self._lib.ma_solver_submit(self._handle, ptr(state_in), ptr(state_out), nsteps))
        # the copies run after this call returns: keep the arrays alive until synchronize()
        self._inflight = getattr(self, "_inflight", []) + [(state_in, state_out)]

    def field(self, which):
        shape = {FIELD_GRADIENT: (self.num_owned_cells, 5, 3), FIELD_LIMITER: (self.num_owned_cells, 5),
                 FIELD_STAGE_PRIMITIVES: (self.num_owned_cells, 5)}[which]
        out = np.empty(shape, dtype=np.float64)
        _abi.check(self._lib.ma_solver_get_field(self._handle, which, out.ctypes.data))
        return out

    def debug_array(self, name, dtype=np.uint8):
        """Test hook (ma_solver_debug_array): one of the solver's device-resident layout arrays as a numpy array."""
        size = C.c_size_t(0)
        _abi.check(self._lib.ma_solver_debug_array(self._handle, name.encode(), None, 0, C.byref(size)))
        buf = np.zeros(size.value, dtype=np.uint8)
        if size.value:
            _abi.check(self._lib.ma_solver_debug_array(self._handle, name.encode(), buf.ctypes.data, size.value, C.byref(size)))
        return buf.view(dtype) if size.value else buf

    @property
    def topology_on_device(self):
        size = C.c_size_t(0)
        _abi.check(self._lib.ma_solver_debug_array(self._handle, b"topology_on_device", None, 0, C.byref(size)))
        return size.value == 1

    def timing(self):
        t = _abi.Timing()
        _abi.check(self._lib.ma_solver_get_timing(self._handle, C.byref(t)))
        return {k: getattr(t, k) for k, _ in t._fields_}

    def reset_timing(self):
        _abi.check(self._lib.ma_solver_reset_timing(self._handle))

    def set_profiling(self, enabled):
        _abi.check(self._lib.ma_solver_set_profiling(self._handle, 1 if enabled else 0))


def write_results(path, mesh, solution, precision=6):
    """`results.<rank>` as Solve() writes it (TimeSolverExplicitRK4.h:514-538)."""
    cmesh, _ = _as_c_mesh(mesh)
    sol = np.ascontiguousarray(solution, dtype=np.float64)
    _abi.check(_abi.load().ma_write_results(os.fsencode(path), C.byref(cmesh), sol.ctypes.data, precision))


# ---- device-function probes (unit parity tests of the physics; they run on the GPU) ----------------------
def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def probe_roe_flux(prim_l, prim_r, normal, tangent, binormal, arith=ARITH_FAST, device=0):
    a = [_c(x) for x in (prim_l, prim_r, normal, tangent, binormal)]
    n = a[0].shape[0]
    out = np.empty((n, 5))
    _abi.check(_abi.load().ma_probe_roe_flux(n, *[x.ctypes.data for x in a], out.ctypes.data, arith, device))
    return out


def probe_viscous_flux(grad, prim, normal, arith=ARITH_FAST, device=0):
    a = [_c(x) for x in (grad, prim, normal)]
    n = a[1].shape[0]
    out = np.empty((n, 5))
    _abi.check(_abi.load().ma_probe_viscous_flux(n, *[x.ctypes.data for x in a], out.ctypes.data, arith, device))
    return out


def probe_primitives(cons, arith=ARITH_FAST, device=0):
    a = _c(cons)
    out = np.empty_like(a)
    _abi.check(_abi.load().ma_probe_primitives(a.shape[0], a.ctypes.data, out.ctypes.data, arith, device))
    return out


def probe_venkat(dumax, dumin, du, deltax3, arith=ARITH_FAST, device=0):
    a = [_c(x) for x in (dumax, dumin, du, deltax3)]
    out = np.empty_like(a[0])
    _abi.check(_abi.load().ma_probe_venkat(a[0].size, *[x.ctypes.data for x in a], out.ctypes.data, arith, device))
    return out


def probe_vanalbada(dumax, dumin, du, arith=ARITH_FAST, device=0):
    a = [_c(x) for x in (dumax, dumin, du)]
    out = np.empty_like(a[0])
    _abi.check(_abi.load().ma_probe_vanalbada(a[0].size, *[x.ctypes.data for x in a], out.ctypes.data, arith, device))
    return out
