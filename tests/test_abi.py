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


def test_header_is_plain_c_and_a_c_caller_links(lib, tmp_path):
    """The boundary is a C ABI: include/miniaero_b200.h compiles as strict C99 (no C++ types, no torch), a caller
    written in C links against the library, and the struct sizes the C compiler sees are the ctypes mirror's.  On a box
    without a GPU the solver constructor must fail with MA_ERR_CUDA and a message, not crash."""
    import subprocess
    from miniaero_b200 import _abi
    import miniaero_b200.build as b
    src = tmp_path / "caller.c"
    src.write_text(r'''
#include <stdio.h>
#include "miniaero_b200.h"
int main(void) {
  ma_options opt;
  ma_solver_config cfg;
  ma_mesh_storage *mesh = NULL;
  ma_solver *solver = NULL;
  int rc;
  ma_options_default(&opt);
  ma_solver_config_default(&cfg);
  opt.nx = 4, opt.ny = 3, opt.nz = 2;
  printf("sizes %d %d %d %d %d\n", (int)sizeof(ma_options), (int)sizeof(ma_faces), (int)sizeof(ma_mesh),
         (int)sizeof(ma_solver_config), (int)sizeof(ma_timing));
  printf("abi %d\n", ma_abi_version());
  if (ma_mesh_generate(&opt, 0, 1, &mesh) != MA_OK) return 2;
  printf("cells %d faces %d\n", ma_mesh_view(mesh)->num_owned_cells, ma_mesh_view(mesh)->internal_faces.nfaces);
  rc = ma_solver_create(ma_mesh_view(mesh), &opt, &cfg, &solver);
  printf("create %d %s\n", rc, rc == MA_OK ? "" : ma_last_error());
  if (rc == MA_OK) ma_solver_destroy(solver);
  ma_mesh_free(mesh);
  return 0;
}
''')
    exe = tmp_path / "caller"
    libdir = os.path.dirname(b.LIB)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe), "-L" + libdir, "-lminiaero_b200", "-Wl,-rpath," + libdir], check=True)
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    out = dict(line.split(" ", 1) for line in p.stdout.strip().splitlines())
    sizes = [int(x) for x in out["sizes"].split()]
    assert sizes == [C.sizeof(_abi.Options), C.sizeof(_abi.Faces), C.sizeof(_abi.Mesh), C.sizeof(_abi.SolverConfig),
                     C.sizeof(_abi.Timing)]
    assert out["abi"].strip() == "3"
    assert out["cells"].strip() == "24 faces 46"   # 4x3x2 cells: 3*3*2 + 4*2*2 + 4*3*1 internal faces
    import torch
    if not torch.cuda.is_available():
        assert out["create"].startswith("-2 ") and "CUDA" in out["create"]   # MA_ERR_CUDA
