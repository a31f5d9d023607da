"""The compiled drop-in: the UNMODIFIED reference's Main.C and mesh generator (Parallel3DMesh, MeshProcessor), built
by oracle/build_ref.sh with the solver class at Main.C:139-141 replaced by include/reference_binding/TimeSolverB200.h
(which calls ma_solver_create / ma_solver_solve of libminiaero_b200.so), run on the reference's three serial
integration tests: miniaero.inp -> results.0 against results.gold at the tolerances of tests/<case>/<case>_test.sh."""
import os
import subprocess

import numpy as np
import pytest

import cases
import parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "miniAero.b200")
HEADERS = [os.path.join(ROOT, "include", "reference_binding", f) for f in ("TimeSolverB200.h", "use_b200_solver.h")]


def _write_inp(path, inp):
    o = cases.opts_kwargs(inp)
    path.write_text("%d\n%r %r %r %r\n%d %d %d\n%d\n%r\n%d\n%d\n%d\n%d\n" % (
        o["problem_type"], o["lx"], o["ly"], o["lz"], o["angle"], o["nx"], o["ny"], o["nz"], o["ntimesteps"], o["dt"],
        1, 100, o["second_order_space"], o["viscous"]))


def test_binding_is_built_against_the_unmodified_reference(lib):
    """CPU: the binding header compiles and links (the binary exists after build()); without a GPU it fails loudly
    through the C ABI's error channel instead of computing anything on the host."""
    assert all(os.path.isfile(h) for h in HEADERS)
    if not os.path.isfile(EXE):
        pytest.skip("oracle/_ref/miniAero.b200 not built (needs /root/reference at build time)")
    ldd = subprocess.run(["ldd", EXE], capture_output=True, text=True).stdout
    assert "libminiaero_b200.so" in ldd and "not found" not in ldd.split("libminiaero_b200.so")[1].splitlines()[0]
    import torch
    if torch.cuda.is_available():
        return
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        from pathlib import Path
        _write_inp(Path(d) / "miniaero.inp", dict(cases.REFERENCE_TESTS["3D_Sod_Serial"][0], ntimesteps=1))
        p = subprocess.run([EXE], cwd=d, capture_output=True, text=True, timeout=300)
        assert p.returncode == 1 and "no CPU fallback" in p.stderr and "TimeSolverB200" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(cases.REFERENCE_TESTS))
@pytest.mark.parametrize("arith", ["fast", "strict"])
def test_reference_main_with_the_b200_solver(lib, tmp_path, name, arith):
    import refrun
    if not os.path.isfile(EXE):
        pytest.skip("oracle/_ref/miniAero.b200 not built")
    inp, rel_tol, floor = cases.REFERENCE_TESTS[name]
    _write_inp(tmp_path / "miniaero.inp", inp)
    p = subprocess.run([EXE], cwd=tmp_path, capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, MINIAERO_B200_ARITH=arith))
    assert p.returncode == 0, p.stdout + p.stderr
    assert "Device Run time" in p.stdout and "Setup time" in p.stdout          # Main.C:127, TimeSolverExplicitRK4.h:495
    res = np.loadtxt(tmp_path / "results.0")
    gold = parity.golden(name)["results_gold"]
    assert res.shape == gold.shape
    assert refrun.numeric_text_diff(res, gold, rel_tol, floor) == 0
    # and against the full-precision reference solution, to the six significant digits results.0 carries
    full = parity.golden(name)["cell_step%d" % inp["ntimesteps"]]
    linf, _ = parity.field_errors(res[:, 3:], full)
    assert linf <= 1e-5
