"""CPU tests of the host side of the boundary: the in-code hex mesh generator (ma_mesh_generate, the
replacement of Parallel3DMesh + MeshProcessor + Face + ElementTopoHexa8) against the arrays the UNMODIFIED
reference builds for the same input (dumped by the oracle's deep_copy hook), and the results.<rank> writer."""
import numpy as np
import pytest

import cases
import parity
import refrun

needs_ref = pytest.mark.skipif(refrun.ref_binary("cell") is None, reason="oracle/_ref not built")

FACE_LABELS = ["face_cell_conn", "cell_flux_index", "coordinates", "face_normal", "face_tangent", "face_binormal"]
# copy order of the reference (Parallel3DMesh.h:365-371) -> position in MeshData::boundary_faces
# (Parallel3DMesh.h:382-396: bottom, top, front, back, right, left)
REF_SET_ORDER = ["internal", "top", "bottom", "right", "left", "front", "back"]
OUR_SET_INDEX = {"bottom": 0, "top": 1, "front": 2, "back": 3, "right": 4, "left": 5}


def _reference_mesh(inp):
    out = refrun.run_reference(dict(inp, ntimesteps=1), kind="cell")
    dumps = out["dumps"]
    sets, i = {}, 0
    for name in REF_SET_ORDER:
        d = {}
        for lab in FACE_LABELS:
            assert dumps[i][0] == lab, (dumps[i][0], lab)
            d[lab] = dumps[i][1]
            i += 1
        assert dumps[i][0] == "sort_order"
        i += 1
        sets[name] = d
    rest = dict((l, a) for l, a in dumps[i:])
    return sets, rest["cell_volumes"], rest["cell_coordinates"]


def _keyed(conn, slot, *arrays):
    """rows keyed by (elem1, slot1): unique per face, independent of the reference's random_shuffle."""
    key = conn[:, 0].astype(np.int64) * 8 + slot[:, 0]
    order = np.argsort(key, kind="stable")
    assert len(np.unique(key)) == len(key)
    return [key[order]] + [a[order] for a in (conn, slot) + arrays]


@needs_ref
@pytest.mark.parametrize("name", ["3D_Sod_Serial", "Ramp_Serial", "FlatPlate_Serial", "ramp_odd"])
def test_generated_mesh_is_bit_identical_to_the_reference(lib, name):
    import miniaero_b200 as ma
    inp = cases.all_cases()[name]
    ref_sets, ref_vol, ref_xyz = _reference_mesh(inp)
    mesh = ma.Parallel3DMesh.from_options(ma.Options(**cases.opts_kwargs(inp))).fillMeshData()
    assert mesh.num_ghosts == 0 and mesh.num_owned_cells == inp["nx"] * inp["ny"] * inp["nz"]
    assert parity.max_ulp(mesh.cell_volumes, ref_vol) == 0
    assert parity.max_ulp(mesh.cell_coordinates, ref_xyz) == 0
    expected_types = {0: ["Tangent", "Tangent", "Tangent", "Tangent", "Extrapolate", "Extrapolate"],
                      1: ["NoSlip", "Extrapolate", "Tangent", "Tangent", "Extrapolate", "Inflow"],
                      2: ["Tangent", "Tangent", "Tangent", "Tangent", "Extrapolate", "Inflow"]}[inp["problem_type"]]
    assert [t for t, _ in mesh.boundary_faces] == expected_types
    for set_name, r in ref_sets.items():
        f = mesh.internal_faces if set_name == "internal" else mesh.boundary_faces[OUR_SET_INDEX[set_name]][1]
        assert f.nfaces_ == r["face_cell_conn"].shape[0], set_name
        ours = _keyed(f.face_cell_conn_, f.cell_flux_index_, f.coordinates_, f.face_normal_, f.face_tangent_,
                      f.face_binormal_)
        theirs = _keyed(r["face_cell_conn"], r["cell_flux_index"], r["coordinates"], r["face_normal"],
                        r["face_tangent"], r["face_binormal"])
        assert np.array_equal(ours[0], theirs[0]), set_name
        if set_name == "internal":
            assert np.array_equal(ours[1], theirs[1]) and np.array_equal(ours[2], theirs[2])
        else:  # elem2 / slot2 of a boundary face are ignored by every functor
            assert np.array_equal(ours[1][:, 0], theirs[1][:, 0]) and np.array_equal(ours[2][:, 0], theirs[2][:, 0])
        for a, b, what in zip(ours[3:], theirs[3:], ["centroid", "normal", "tangent", "binormal"]):
            assert parity.max_ulp(a, b) == 0, (set_name, what)


def test_slot_convention_and_closed_cells(lib):
    """slot = local hex face 0:-y 1:+x 2:+y 3:-x 4:-z 5:+z; cell id = i*ny*nz + j*nz + k;
    elem1 = lower cell id; area vectors of every cell sum to ~0 (closed control volumes)."""
    import miniaero_b200 as ma
    nx, ny, nz = 5, 4, 3
    mesh = ma.Parallel3DMesh(nx, ny, nz, 1.0, 0.8, 0.6, 2, angle=20.0).fillMeshData()
    f = mesh.internal_faces
    assert f.nfaces_ == (nx - 1) * ny * nz + nx * (ny - 1) * nz + nx * ny * (nz - 1)
    l, r = f.face_cell_conn_[:, 0], f.face_cell_conn_[:, 1]
    assert (l < r).all()
    d = r - l
    sl, sr = f.cell_flux_index_[:, 0], f.cell_flux_index_[:, 1]
    assert ((d == ny * nz) == ((sl == 1) & (sr == 3))).all()
    assert ((d == nz) == ((sl == 2) & (sr == 0))).all()
    assert ((d == 1) == ((sl == 5) & (sr == 4))).all()
    acc = np.zeros((mesh.num_owned_cells, 3))
    np.add.at(acc, l, f.face_normal_)
    np.add.at(acc, r, -f.face_normal_)
    for _, b in mesh.boundary_faces:
        np.add.at(acc, b.face_cell_conn_[:, 0], b.face_normal_)
    scale = np.abs(f.face_normal_).max()
    assert np.abs(acc).max() < 1e-12 * scale
    # unit tangent, binormal = a x t (Face.C:81-96)
    assert np.allclose(np.linalg.norm(f.face_tangent_, axis=1), 1.0, atol=1e-14)
    assert np.allclose(f.face_binormal_, np.cross(f.face_normal_, f.face_tangent_), atol=1e-18)
    assert (mesh.cell_volumes > 0).all()


@needs_ref
def test_results_writer_matches_the_reference_text(lib, tmp_path):
    """ma_write_results reproduces results.<rank> (TimeSolverExplicitRK4.h:514-538) character for character
    when given the reference's own solution."""
    import miniaero_b200 as ma
    inp = dict(cases.EXTRA["ramp_odd"], ntimesteps=2)
    out = refrun.run_reference(inp, kind="cell")
    ref_text = open(out["workdir"] + "/results.0").read()
    mesh = ma.Parallel3DMesh.from_options(ma.Options(**cases.opts_kwargs(inp))).fillMeshData()
    ma.write_results(str(tmp_path / "results.0"), mesh, refrun.solution_from_dumps(out["dumps"]))
    assert open(tmp_path / "results.0").read() == ref_text


@pytest.mark.skipif(not __import__("os").path.isfile(__import__("os").path.join(refrun.REF_DIR, "miniAero.cell.mpi")),
                    reason="oracle/_ref/miniAero.cell.mpi not built")
@pytest.mark.parametrize("name,nranks", [("sod_o2_visc", 2), ("sod_o2_visc", 8), ("FlatPlate_Parallel", 8), ("ramp_o2_visc", 4)])
def test_block_meshes_match_the_reference_mpi_build(lib, name, nranks):
    """Per rank: owned + ghost cell order, centroids (bit for bit) and owned volumes equal what the reference's
    WITH_MPI mesh setup (Parallel3DMesh.C:98-174, 306-431) hands its solver."""
    import miniaero_b200 as ma
    inp = cases.PARALLEL[name][0]
    out = refrun.run_reference_parallel(dict(inp, ntimesteps=1), nranks, labels="cell_coordinates,cell_volumes")
    for r, o in enumerate(out):
        d = dict(o["dumps"])
        mesh = ma.Parallel3DMesh.from_options(ma.Options(**cases.opts_kwargs(inp)), r, nranks).fillMeshData()
        n = mesh.num_owned_cells
        assert d["cell_coordinates"].shape[0] == n + mesh.num_ghosts, (r, d["cell_coordinates"].shape, n, mesh.num_ghosts)
        assert o["results"].shape[0] == n
        assert parity.max_ulp(mesh.cell_coordinates, d["cell_coordinates"]) == 0, r
        assert parity.max_ulp(mesh.cell_volumes[:n], d["cell_volumes"][:n]) == 0, r


def test_partitioned_norms_equal_the_global_field_norms():
    """parity.combine_parts over the blocks of a partitioned field == parity.field_errors of the whole field; a block
    whose momentum is ~0 (the wave has not arrived) is measured on the global momentum scale, not its own."""
    import parity
    rng = np.random.default_rng(7)
    ref = rng.normal(size=(96, 5))
    ref[:, 4] += 1e5
    ref[:32, 1:4] *= 1e-9          # a block that is still (almost) at rest
    test = ref + 1e-13 * rng.normal(size=ref.shape)
    whole = parity.field_errors(test, ref)
    parts = [parity.field_error_parts(test[a:b], ref[a:b]) for a, b in ((0, 32), (32, 64), (64, 96))]
    combined = parity.combine_parts(parts)
    assert np.allclose(whole, combined, rtol=1e-12)
    assert parity.field_errors(test[:32], ref[:32])[0] > 1e3 * combined[0]   # the per-block norm would cry wolf
