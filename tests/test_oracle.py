"""CPU tests that PIN the oracle: the unmodified reference built in oracle/_ref (oracle/build_ref.sh) must
(a) pass the reference's own integration tests — its 6-digit gold files, re-encoded in tests/golden/*.npz —
at the tolerances of tests/<case>/<case>_test.sh, and (b) reproduce, bit for bit, the full-precision golden
vectors the GPU parity tests compare against (so those vectors provably come from this oracle)."""
import os
import sys

import numpy as np
import pytest

import cases
import parity
import refrun

pytestmark = pytest.mark.skipif(refrun.ref_binary("cell") is None,
                                reason="oracle/_ref not built (needs /root/reference once; the binaries travel)")


@pytest.mark.parametrize("name", sorted(cases.REFERENCE_TESTS))
def test_oracle_passes_the_reference_gold_files(name):
    inp, rel_tol, floor = cases.REFERENCE_TESTS[name]
    out = refrun.run_reference(inp, kind="cell")
    g = parity.golden(name)
    assert out["results"].shape == g["results_gold"].shape
    assert refrun.numeric_text_diff(out["results"], g["results_gold"], rel_tol, floor) == 0
    full = refrun.solution_from_dumps(out["dumps"])
    assert parity.max_ulp(full, g["cell_step%d" % inp["ntimesteps"]]) == 0


@pytest.mark.parametrize("name", sorted(cases.EXTRA))
def test_golden_vectors_come_from_this_oracle(name):
    inp = cases.EXTRA[name]
    g = parity.golden(name)
    for n in (1, 2):
        out = refrun.run_reference(dict(inp, ntimesteps=n), kind="cell")
        assert parity.max_ulp(refrun.solution_from_dumps(out["dumps"]), g["cell_step%d" % n]) == 0


def test_openmp_build_is_bit_identical_to_serial():
    """-DCELL_FLUX is deterministic and thread-count independent: the CPU-baseline binary (OpenMP static loop)
    gives the serial oracle's bits."""
    inp = dict(cases.EXTRA["sod_o2_visc"], ntimesteps=2)
    a = refrun.solution_from_dumps(refrun.run_reference(inp, kind="cell")["dumps"])
    b = refrun.solution_from_dumps(refrun.run_reference(inp, kind="cell", omp=True, threads=4)["dumps"])
    assert parity.max_ulp(a, b) == 0
    assert parity.max_ulp(a, parity.golden("sod_o2_visc")["cell_step2"]) == 0


def test_reference_noise_floor_between_its_two_builds():
    """The Makefile-default -DATOMICS_FLUX build differs from -DCELL_FLUX only by summation order: the
    distance after 100 steps is the reference's own noise floor and sits inside the north-star tolerance."""
    for name in ("3D_Sod_Serial", "sod_o2_visc", "flatplate_o1"):
        g = parity.golden(name)
        linf, l2 = parity.field_errors(g["atomics_step100"], g["cell_step100"])
        assert linf < parity.TOL_100_STEPS and l2 < parity.TOL_100_STEPS, (name, linf, l2)


def test_launch_log_gives_per_step_times(tmp_path, monkeypatch):
    """bench.py's CPU baseline reads per-time-step timestamps from the stand-in's launch log."""
    log = tmp_path / "launch.log"
    monkeypatch.setenv("MINIAERO_LAUNCH_LOG", str(log))
    refrun.run_reference(dict(cases.EXTRA["sod_o2"], ntimesteps=3, output_results=0), kind="cell", omp=True, threads=2,
                         dump=False)
    lines = log.read_text().split("\n")
    ends = [float(l.split()[1]) for l in lines if l.startswith("step_end")]
    assert len(ends) == 4 and ends == sorted(ends)
    assert any(l.startswith("functor") and "compute_face_flux" in l for l in lines)


def test_numeric_text_diff_semantics():
    a = np.array([[1.0, 1e-9], [2.0, 5.0]])
    assert refrun.numeric_text_diff(a, a) == 0
    b = a.copy(); b[0, 1] = 5e-9          # below the floor: ignored
    assert refrun.numeric_text_diff(a, b) == 0
    b[1, 1] = 5.02                        # 0.4 % > 1e-3
    assert refrun.numeric_text_diff(a, b) == 1
    assert refrun.numeric_text_diff(a, b, rel_tol=1e-2) == 0


# ---- the parallel oracle: the reference's WITH_MPI build over the file-based MPI stand-in -------------------
needs_mpi_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(refrun.REF_DIR, "miniAero.cell.mpi")),
                                   reason="oracle/_ref/miniAero.cell.mpi not built")


@needs_mpi_ref
@pytest.mark.parametrize("name,nranks", [("3D_Sod_Parallel", 4), ("FlatPlate_Parallel", 8)])
def test_parallel_oracle_passes_the_reference_parallel_gold_files(name, nranks):
    """tests/3D_Sod_Parallel (mpirun -np 4) and tests/FlatPlate_Parallel (-np 8): per-rank results.<rank> against
    results.<rank>.gold at the scripts' tolerances, and the committed full-precision per-rank vectors are its bits."""
    inp, _, _, rel_tol, floor = cases.PARALLEL[name]
    g = np.load(os.path.join(parity.GOLDEN, "par_%s_%d.npz" % (name, nranks)))
    out = refrun.run_reference_parallel(inp, nranks)
    for r, o in enumerate(out):
        assert refrun.numeric_text_diff(o["results"], g["r%d_gold" % r], rel_tol, floor) == 0, (name, r)
        nown = o["results"].shape[0]
        full = refrun.solution_from_dumps(o["dumps"])[:nown]
        assert parity.max_ulp(full, g["r%d_step%d" % (r, inp["ntimesteps"])]) == 0


@needs_mpi_ref
def test_parallel_oracle_agrees_with_serial_oracle_to_roundoff():
    """Block seams flip the left/right roles of some faces (owned cell is always elem1), and the Roe flux is not
    bitwise antisymmetric, so per-rank results differ from the single-domain run by roundoff only."""
    name, nranks = "sod_o2_visc", 4
    inp = cases.PARALLEL[name][0]
    g = np.load(os.path.join(parity.GOLDEN, "par_%s_%d.npz" % (name, nranks)))
    serial = parity.golden(name)["cell_step100"]
    import miniaero_b200 as ma
    worst = 0.0
    for r in range(nranks):
        mesh = ma.Parallel3DMesh.from_options(ma.Options(**cases.opts_kwargs(inp)), r, nranks).fillMeshData()
        gids = mesh.global_ids[:mesh.num_owned_cells]
        linf, l2 = parity.field_errors(g["r%d_step100" % r], serial[gids])
        worst = max(worst, linf, l2)
    assert worst < parity.TOL_100_STEPS, worst


def test_unit_golden_vectors_come_from_the_reference_functions():
    """tests/golden/unit_functions.npz = tests/unit_inputs.py pushed through the reference's own device functions
    (oracle/_ref/unit_oracle: Roe_Flux.h, Viscous_Flux.h, GasModel.h, VenkatLimiter.h, VanAlbadaLimiter.h)."""
    import unit_inputs
    g = parity.golden("unit_functions")
    for fn, rows in unit_inputs.make_inputs().items():
        assert parity.max_ulp(rows, g[fn + "_in"]) == 0, fn
        assert parity.max_ulp(refrun.unit_oracle(fn, rows), g[fn + "_out"]) == 0, fn
    # the sample reaches every branch: Venkat's |du| <= 1e-40 -> 1, Van Albada's clamp, the Roe eigenvalue fix
    assert (g["venkat_out"] == 1.0).sum() >= 64 and g["venkat_out"].max() > 1.0
    assert (g["vanalbada_out"] == 1.0).any() and (g["vanalbada_out"] == 0.0).any()
    x = g["roe_in"]
    a = x[:, 10:13]
    un = (x[:, 1:4] * a).sum(axis=1) / np.linalg.norm(a, axis=1)
    c = np.sqrt(1.4 * 287.05 * x[:, 4])
    assert (np.abs(un) < 0.05 * c).any() and (np.abs(np.abs(un) - c) < 0.05 * c).any()


def test_fullsize_column_golden_comes_from_this_oracle():
    """tests/golden/sod_fullsize_column.npz: the reference's x-line for the cell size of the 67 M-cell benchmark mesh,
    and its own 16 lines agree to roundoff (the property tests/test_gpu_fullsize.py relies on)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_fullsize as mk
    g = parity.golden("sod_fullsize_column")
    out = refrun.run_reference(dict(mk.INP, ntimesteps=2), kind="cell")
    sol = refrun.solution_from_dumps(out["dumps"]).reshape(mk.NX, 4, 4, 5)
    assert parity.max_ulp(sol[:, 1, 1, :], g["line_step2"]) == 0
    spread = np.abs(sol - sol[:, 1:2, 1:2, :]).max(axis=(0, 1, 2))
    scale = np.abs(sol).max(axis=(0, 1, 2))
    scale[1:4] = np.sqrt((sol[..., 1:4] ** 2).sum(axis=-1)).max()
    assert (spread <= 1e-14 * scale).all()


@pytest.mark.parametrize("name", ["sod_o2", "ramp_odd"])
def test_vanalbada_golden_vectors_come_from_the_redirected_reference(name):
    """tests/golden/vanalbada_*.npz = the unmodified reference sources with the stencil limiter class redirected to
    VanAlbadaLimiter (oracle/vanalbada_swap.h); a different scheme from the Venkatakrishnan default after a step."""
    g = parity.golden("vanalbada_" + name)
    out = refrun.run_reference(dict(cases.EXTRA[name], ntimesteps=2), kind="cell.vanalbada")
    assert parity.max_ulp(refrun.solution_from_dumps(out["dumps"]), g["cell_step2"]) == 0
    assert parity.max_ulp(g["cell_step2"], parity.golden(name)["cell_step2"]) != 0


def test_initial_condition_golden_comes_from_this_oracle():
    g = parity.golden("initial_conditions")
    for name in g.files:
        out = refrun.run_reference(dict(cases.EXTRA[name], ntimesteps=0), kind="cell")
        assert parity.max_ulp(refrun.solution_from_dumps(out["dumps"]), g[name]) == 0
