"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/miniaero_b200.h
declares, and its host-only entry points (options, error channel) behave; no compute call needs a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "miniaero_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ma_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from miniaero_b200 import _abi
    declared = _declared_functions()
    assert len(declared) >= 30
    assert sorted(_abi.SYMBOLS) == declared, "ctypes table and header disagree"
    for name in declared:
        assert getattr(lib, name) is not None


def test_abi_version_and_struct_sizes(lib):
    from miniaero_b200 import _abi
    assert lib.ma_abi_version() == 3
    # plain-C layout checks (no torch / C++ types cross the boundary)
    assert C.sizeof(_abi.Options) == 80
    assert C.sizeof(_abi.Faces) == 56
    assert _abi.Mesh.boundary_faces.offset % 8 == 0


def test_options_default_and_read(lib, tmp_path):
    import miniaero_b200 as ma
    o = ma.Options()
    # Options.h:59-69 defaults
    assert (o.problem_type, o.nx, o.ny, o.nz, o.ntimesteps) == (0, 10, 10, 10, 1)
    assert o.second_order_space == 0 and o.viscous == 0
    p = tmp_path / "miniaero.inp"
    # the reference's tests/FlatPlate_Serial/miniaero.inp layout (Options.h:91-99)
    p.write_text("1\n2.0 0.002 1.0 0.0\n16 32 2\n400\n3e-8\n1\n100\n1\n1\n")
    o.read_options_file(str(p))
    assert (o.problem_type, o.lx, o.ly, o.lz, o.angle) == (1, 2.0, 0.002, 1.0, 0.0)
    assert (o.nx, o.ny, o.nz, o.ntimesteps, o.dt) == (16, 32, 2, 400, 3e-8)
    assert (o.output_results, o.output_frequency, o.second_order_space, o.viscous) == (1, 100, 1, 1)


def test_options_errors_are_reported_not_ignored(lib, tmp_path):
    import miniaero_b200 as ma
    with pytest.raises(ma.MiniAeroError, match="does_not_exist"):
        ma.Options().read_options_file(str(tmp_path / "does_not_exist.inp"))
    short = tmp_path / "short.inp"
    short.write_text("0\n1.0 1.0\n")
    with pytest.raises(ma.MiniAeroError):
        ma.Options().read_options_file(str(short))


def test_solver_refuses_to_run_without_a_gpu(lib):
    """No CPU fallback: on a box without a CUDA device the constructor fails loudly with MA_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import miniaero_b200 as ma
    opt = ma.Options(nx=4, ny=4, nz=4, dt=1e-6)
    mesh = ma.Parallel3DMesh.from_options(opt).fillMeshData()
    with pytest.raises(ma.MiniAeroError, match="error -2"):
        ma.TimeSolverExplicitRK4(mesh, opt)
    with pytest.raises(ma.MiniAeroError, match="error -2"):
        ma.probe_primitives([[1.0, 0, 0, 0, 2.5e5]])


def test_invalid_arguments(lib):
    import miniaero_b200 as ma
    with pytest.raises(ma.MiniAeroError, match="error -1"):
        ma.Parallel3DMesh(0, 4, 4, 1.0, 1.0, 1.0, 0).fillMeshData()
    with pytest.raises(ma.MiniAeroError, match="error -1"):
        ma.Parallel3DMesh(4, 4, 4, 1.0, 1.0, 1.0, 0, rank=0, num_ranks=3).fillMeshData()  # ranks must be 2^k
    with pytest.raises(AttributeError):
        ma.Options(no_such_field=1)
