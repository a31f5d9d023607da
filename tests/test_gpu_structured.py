"""ma_solver_create_structured: the block's device layout built straight from (i, j, k) with
Face.C / ElementTopoHexa8 geometry evaluated ON THE DEVICE gives the solver that the reference-format mesh path
(ma_mesh_generate -> ma_solver_create) gives — identical bits after time steps, in both arithmetic modes, and
therefore the reference's results (STRICT: bit for bit against tests/golden)."""
import numpy as np
import pytest

import cases
import parity

pytestmark = pytest.mark.gpu


def _both(inp, nsteps, arith, **kw):
    import miniaero_b200 as ma
    opt = ma.Options(**cases.opts_kwargs(dict(inp, ntimesteps=nsteps)))
    mesh = ma.Parallel3DMesh.from_options(opt).fillMeshData()
    a = ma.TimeSolverExplicitRK4(mesh, opt, arith=arith, **kw)
    b = ma.TimeSolverExplicitRK4.from_options(opt, arith=arith, **kw)
    out = []
    for s in (a, b):
        s.initialize()
        s.step(nsteps)
        out.append(s.solution())
    assert b.num_owned_cells == mesh.num_owned_cells
    return out


@pytest.mark.parametrize("name", sorted(cases.all_cases()))
def test_structured_constructor_gives_the_same_bits(lib, name):
    import miniaero_b200 as ma
    inp = cases.all_cases()[name]
    ref = parity.golden(name)["cell_step2"]
    mesh_path, structured = _both(inp, 2, ma.ARITH_STRICT)
    assert parity.max_ulp(structured, mesh_path) == 0
    assert parity.max_ulp(structured, ref) == 0          # the reference's -DCELL_FLUX build, bit for bit
    mesh_path, structured = _both(inp, 2, ma.ARITH_FAST)
    assert parity.max_ulp(structured, mesh_path) == 0


def test_structured_constructor_other_tiles_and_fields(lib):
    """Ragged tiles, a sheared mesh (the two paths cut different tiles there: results do not depend on the tiling),
    and the intermediate fields."""
    import miniaero_b200 as ma
    inp = cases.EXTRA["ramp_odd"]
    for tile in ((4, 4, 4), (3, 5, 2), (8, 8, 8)):
        a, b = _both(inp, 2, ma.ARITH_FAST, tile_dims=tile)
        assert parity.max_ulp(a, b) == 0
    opt = ma.Options(**cases.opts_kwargs(dict(inp, ntimesteps=1)))
    mesh = ma.Parallel3DMesh.from_options(opt).fillMeshData()
    sa = ma.TimeSolverExplicitRK4(mesh, opt, arith=ma.ARITH_STRICT)
    sb = ma.TimeSolverExplicitRK4.from_options(opt, arith=ma.ARITH_STRICT)
    for s in (sa, sb):
        s.initialize()
        s.step(1)
    for which in (ma.FIELD_GRADIENT, ma.FIELD_LIMITER, ma.FIELD_STAGE_PRIMITIVES):
        assert parity.max_ulp(sa.field(which), sb.field(which)) == 0


def test_structured_constructor_errors(lib):
    import miniaero_b200 as ma
    opt = ma.Options(**cases.opts_kwargs(dict(cases.EXTRA["sod_o2"], ntimesteps=1)))
    with pytest.raises(ma.MiniAeroError):
        ma.TimeSolverExplicitRK4.from_options(opt, rank=0, nranks=3)        # not a power of two (Parallel3DMesh.C:262)
    with pytest.raises(ma.MiniAeroError):
        ma.TimeSolverExplicitRK4.from_options(opt, rank=0, nranks=2)        # two ranks need a communicator
