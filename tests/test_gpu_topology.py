"""Device-side topology: ma_solver_create_structured builds the O(cells) layout arrays — cell
renumbering, slot maps, tile-local connectivity, outside-cell and publish lists, and through them the device-evaluated
geometry — on the GPU from per-pattern templates (topology_kernels.cu).  They must equal, byte for byte, what the host
builder (layout.cpp, MINIAERO_DEVICE_TOPOLOGY=0) uploads; tools/topology_compare.cpp checks the same logic on the CPU."""
import os
import subprocess

import numpy as np
import pytest

import cases
import parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARRAYS = ["tiles", "slot_face", "slot_nbr", "face_lr", "tile_halo", "tile_pub", "old2new", "face_geom", "cell_xyz", "cell_vol"]


@pytest.fixture(scope="module")
def comparer(tmp_path_factory):
    from miniaero_b200 import build as b
    b.build()
    exe = str(tmp_path_factory.mktemp("topo") / "topology_compare")
    objs = [os.path.join(b.BUILD, n) for n in ("layout.o", "host_mesh.o", "host_common.o")]
    subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"),
                    "-I" + b.CSRC, "-I" + b._cuda_include(), os.path.join(ROOT, "tools", "topology_compare.cpp")] + objs +
                   ["-o", exe], check=True)
    return exe


@pytest.mark.parametrize("args", [
    (64, 8, 8, 0, 0, 1), (64, 8, 8, 0, 0, 1, 4, 4, 8, 1), (37, 21, 13, 1, 0, 1, 4, 4, 8, 1), (37, 21, 13, 2, 0, 1, 3, 5, 2, 1),
    (64, 32, 32, 0, 1, 4, 4, 4, 8, 1), (64, 32, 32, 1, 3, 8, 4, 4, 8, 1), (64, 32, 32, 1, 7, 8, 4, 4, 8, 0), (16, 9, 7, 0, 0, 2, 4, 4, 4, 0),
    (16, 9, 7, 0, 1, 2, 4, 4, 4, 1), (5, 3, 2, 0, 0, 1, 8, 8, 8, 1), (128, 4, 4, 0, 0, 1, 4, 4, 8, 1), (9, 9, 9, 2, 0, 1, 16, 2, 2, 1),
    (96, 64, 64, 0, 5, 8, 4, 4, 8, 1)])
def test_stamping_logic_matches_the_host_builder(comparer, args):
    """No GPU: build_topology_plan + topology_stamp.h (the functions the CUDA kernels run) against the host builder:
    NX NY NZ problem_type rank nranks [tile dims] [shared cut faces]."""
    p = subprocess.run([comparer] + [str(a) for a in args], capture_output=True, text=True)
    assert p.returncode == 0 and "differences: 0" in p.stdout, p.stdout + p.stderr


def _solver(inp, env, rank=0, nranks=1, **kw):
    import miniaero_b200 as ma
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        opt = ma.Options(**cases.opts_kwargs(dict(inp, ntimesteps=3)))
        return ma.TimeSolverExplicitRK4.from_options(opt, rank, nranks, **kw)
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", [("sod_o2_visc", {}), ("sod_o2_visc", {"share_cut_faces": 1}), ("ramp_odd", {"share_cut_faces": 1}),
                                     ("FlatPlate_Serial", {"share_cut_faces": 1, "tile_dims": (4, 4, 4)}),
                                     ("ramp_odd", {"tile_dims": (3, 5, 2)}), ("Ramp_Serial", {"share_cut_faces": -1})])
def test_device_built_arrays_equal_the_host_built_ones(lib, name, kw):
    inp = cases.all_cases()[name]
    dev = _solver(inp, {}, **kw)
    host = _solver(inp, {"MINIAERO_DEVICE_TOPOLOGY": "0"}, **kw)
    assert dev.topology_on_device and not host.topology_on_device
    for arr in ARRAYS:
        a, b = dev.debug_array(arr), host.debug_array(arr)
        assert a.size == b.size and np.array_equal(a, b), arr
    for s in (dev, host):
        s.initialize()
        s.step(3)
    assert parity.max_ulp(dev.solution(), host.solution()) == 0


@pytest.mark.gpu
def test_device_topology_of_a_large_block(lib):
    """A block of 2 M cells with every kind of brick (ragged edges in all three directions), shared cut faces on by
    default: same arrays, and the set-up reports where the topology was built."""
    inp = dict(cases.EXTRA["sod_o2_visc"], nx=203, ny=101, nz=99)
    dev = _solver(inp, {})
    host = _solver(inp, {"MINIAERO_DEVICE_TOPOLOGY": "0"})
    assert dev.topology_on_device and not host.topology_on_device
    for arr in ARRAYS:
        assert np.array_equal(dev.debug_array(arr), host.debug_array(arr)), arr
