"""Host-side tile layout (miniaero_b200/csrc/layout.cpp) without a GPU: tools/layout_check.cpp rebuilds the layout of
in-code meshes and checks the invariants the kernels rely on (tile-local connectivity, outside-cell lists, slot maps)
and the shared-memory bank model of the staged flux kernel."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    from miniaero_b200 import build as b
    b.build()
    exe = str(tmp_path_factory.mktemp("layout") / "layout_check")
    objs = [os.path.join(b.BUILD, n) for n in ("layout.o", "host_mesh.o", "host_common.o")]
    subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-I" + os.path.join(ROOT, "include"), "-I" + b.CSRC,
                    "-I" + b._cuda_include(), os.path.join(ROOT, "tools", "layout_check.cpp")] + objs + ["-o", exe], check=True)
    return exe


def _run(exe, args, env=None):
    p = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True,
                       env=dict(os.environ, **(env or {})))
    assert p.returncode == 0, p.stdout + p.stderr
    assert "invariant violations: 0" in p.stdout
    return p.stdout


@pytest.mark.parametrize("args", [(64, 8, 8), (37, 21, 13), (37, 21, 13, 4, 4, 4), (16, 9, 7, 3, 5, 2), (5, 3, 2, 8, 8, 8),
                                  (128, 4, 4), (9, 9, 9, 16, 2, 2)])
def test_layout_invariants(checker, args):
    _run(checker, args)


@pytest.mark.parametrize("share", ["", "1"])
@pytest.mark.parametrize("args", [(37, 21, 13), (16, 9, 7, 3, 5, 2), (5, 3, 2, 8, 8, 8)])
def test_layout_does_not_depend_on_the_order_of_the_face_list(checker, args, share):
    """The reference's mesh generator hands over its internal faces in a random order (Parallel3DMesh.h:362, an unseeded
    std::random_shuffle): the device layout is a function of the cells' connectivity, not of where a face sits in the
    caller's list — same bytes for the creation order and for two shuffles of it."""
    sums = set()
    for seed in (None, "1", "2"):
        env = {"MINIAERO_CHECK_SHARE": "1"} if share else {}
        if seed:
            env["MINIAERO_CHECK_SHUFFLE"] = seed
        out = _run(checker, args, env)
        sums.add(re.search(r"layout checksum: ([0-9a-f]{16})", out).group(1))
    assert len(sums) == 1, sums


@pytest.mark.parametrize("how,message", [("conn", "out of range"), ("slot", "out of range"),
                                         ("missing", "2 (cell, slot) pairs of owned cells have no face"),
                                         ("duplicate", "2 (cell, slot) pairs of owned cells have no face")])
def test_malformed_meshes_are_refused(checker, how, message):
    """ma_solver_create's mesh checks (layout.cpp, ArrayAccess): a face naming a cell that does not exist, a slot
    outside 0..5, a hex cell with fewer than six faces — each is an error with a text, not a crash or a wrong layout."""
    p = subprocess.run([checker, "9", "7", "5"], capture_output=True, text=True,
                       env=dict(os.environ, MINIAERO_CHECK_CORRUPT=how))
    assert p.returncode == 1 and p.stdout.startswith("layout: mesh: ") and message in p.stdout, p.stdout + p.stderr


@pytest.mark.parametrize("share", ["", "1"])
@pytest.mark.parametrize("args", [(37, 21, 13), (64, 32, 32), (16, 9, 7, 3, 5, 2)])
def test_layout_does_not_depend_on_the_thread_count(checker, args, share):
    """The host builder runs its O(cells) loops in parallel, first-touches its arrays from all threads and shares face
    orders between tiles through a cache filled concurrently: the arrays it hands to the device must be the same bytes
    whatever the number of threads."""
    sums = set()
    for threads in ("1", "3", "8"):
        env = {"OMP_NUM_THREADS": threads}
        if share:
            env["MINIAERO_CHECK_SHARE"] = "1"
        out = _run(checker, args, env)
        sums.add(re.search(r"layout checksum: ([0-9a-f]{16})", out).group(1))
    assert len(sums) == 1, sums


@pytest.mark.parametrize("args", [(64, 8, 8), (37, 21, 13), (37, 21, 13, 4, 4, 4), (16, 9, 7, 3, 5, 2), (5, 3, 2, 8, 8, 8),
                                  (128, 4, 4), (9, 9, 9, 16, 2, 2), (64, 32, 32)])
def test_shared_cut_faces(checker, args):
    """Shared cut faces (layout.h): every face between two tiles is evaluated by exactly one of them — the one of the
    earlier flux launch — and published into the other's import slot; a brick tiling needs no face twice."""
    out = _run(checker, args, {"MINIAERO_CHECK_SHARE": "1"})
    m = re.search(r"shared cut faces: (\d+) evaluations of (\d+) internal faces \((\d+) evaluated twice\), (\d+) imports", out)
    assert m and int(m.group(1)) == int(m.group(2)) and int(m.group(3)) == 0


@pytest.mark.parametrize("env", [{"MINIAERO_FACE_ORDER": "slot"}, {"MINIAERO_FACE_ORDER": "cell"},
                                 {"MINIAERO_CELL_SWIZZLE": "0"}, {"MINIAERO_TILE_ORDER": "linear"}])
def test_layout_invariants_under_every_knob(checker, env):
    _run(checker, (24, 16, 16), env)


def test_default_order_is_bank_friendly(checker):
    """4x4x8 bricks: the swizzled cell order plus the half-warp packing of the face lists keep the record reads of
    the flux kernel's face loop within 1.25 wavefronts per ideal wavefront (1.67 for the plain slot order)."""
    out = _run(checker, (64, 32, 32))
    ratio = float(re.search(r"phase 1 record-read wavefronts / ideal: ([0-9.]+)", out).group(1))
    assert ratio < 1.25, out
    plain = _run(checker, (64, 32, 32), {"MINIAERO_FACE_ORDER": "slot", "MINIAERO_CELL_SWIZZLE": "0"})
    assert float(re.search(r"phase 1 record-read wavefronts / ideal: ([0-9.]+)", plain).group(1)) > ratio


@pytest.fixture(scope="module")
def comparer(tmp_path_factory):
    from miniaero_b200 import build as b
    b.build()
    exe = str(tmp_path_factory.mktemp("layout") / "layout_compare")
    objs = [os.path.join(b.BUILD, n) for n in ("layout.o", "host_mesh.o", "host_common.o")]
    subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"),
                    "-I" + b.CSRC, os.path.join(ROOT, "tools", "layout_compare.cpp")] + objs + ["-o", exe], check=True)
    return exe


# NX NY NZ PROBLEM_TYPE ANGLE RANK NRANKS [tile] [strict]
@pytest.mark.parametrize("args", [(37, 21, 13, 0, 0, 0, 1), (16, 32, 2, 1, 0, 0, 1), (64, 32, 2, 2, 30, 0, 1),
                                  (32, 16, 8, 2, 30, 1, 4), (32, 64, 2, 1, 0, 5, 8), (13, 7, 5, 2, 17, 0, 1, 3, 5, 2),
                                  (24, 16, 16, 0, 0, 1, 2, 8, 8, 8, 1), (64, 8, 8, 0, 0, 1, 2), (128, 4, 4, 0, 0, 3, 4)])
def test_structured_layout_is_the_array_layout(comparer, args):
    """build_layout_structured() (layout from (i, j, k), nothing materialised) == build_layout(ma_mesh_generate()),
    every array bit for bit, and the deferred-geometry face codes re-evaluate to the same geometry."""
    p = subprocess.run([comparer] + [str(a) for a in args], capture_output=True, text=True)
    assert p.returncode == 0 and "differences: 0" in p.stdout, p.stdout + p.stderr
