"""Host driver (miniaero_b200/miniaero, the reference's Main.C shape) and the Mantevo YAML report
(YAML_Doc.C:27-67 / YAML_Element.C:97-104 grammar)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import cases
import parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "miniaero_b200", "miniaero")


def _write_inp(path, inp):
    """the nine values of miniaero.inp in file order (Options.h:91-99)"""
    o = cases.opts_kwargs(inp)
    path.write_text("%d\n%r %r %r %r\n%d %d %d\n%d\n%r\n%d\n%d\n%d\n%d\n" % (
        o["problem_type"], o["lx"], o["ly"], o["lz"], o["angle"], o["nx"], o["ny"], o["nz"], o["ntimesteps"], o["dt"],
        o.get("output_results", 1), 100, o["second_order_space"], o["viscous"]))


def test_yaml_report_grammar(lib, tmp_path):
    from miniaero_b200 import _abi
    opt = _abi.Options()
    lib.ma_options_default(C.byref(opt))
    opt.nx, opt.ny, opt.nz, opt.ntimesteps, opt.second_order_space, opt.viscous = 512, 512, 256, 10, 1, 1
    tm = _abi.Timing()
    tm.step_seconds, tm.steps, tm.cell_updates, tm.num_tiles = 0.7, 10, 10 * 512 * 512 * 256, 524288
    rep = _abi.Report()
    rep.options, rep.timing = C.pointer(opt), C.pointer(tm)
    rep.num_ranks, rep.global_cells = 1, 512 * 512 * 256
    rep.blocks[0] = rep.blocks[1] = rep.blocks[2] = 1
    rep.setup_seconds, rep.run_seconds, rep.total_seconds, rep.hbm_peak_gbs = 10.0, 1.0, 11.5, 6551.4
    rep.device_name = b"test device"
    out = C.create_string_buffer(512)
    _abi.check(lib.ma_write_yaml_report(C.byref(rep), str(tmp_path).encode(), out, 512))
    path = out.value.decode()
    # <name>-<version>_<YYYY:MM:DD-HH:MM:SS>.yaml (YAML_Doc.C:41-49)
    assert re.fullmatch(r".*/miniAero-b200-1\.0_\d{4}:\d\d:\d\d-\d\d:\d\d:\d\d\.yaml", path)
    lines = open(path).read().splitlines()
    assert lines[0] == "Mini-Application Name: miniAero-b200"
    assert lines[1] == "Mini-Application Version: 1.0"
    # "key: value", children indented two spaces per level (YAML_Element.C:97-104)
    assert all(re.fullmatch(r"(  )*[^:]+: .*", l) or l.endswith(": ") for l in lines)
    kv = dict(l.strip().split(": ", 1) for l in lines if not l.endswith(": "))
    assert kv["global cells"] == str(512 * 512 * 256) and kv["device"] == "test device"
    cups = float(kv["cell-updates per second"])
    assert abs(cups - 512 * 512 * 256 * 10 / 0.7) / cups < 1e-5
    assert abs(float(kv["fraction of HBM roofline"]) - cups * 4648 / 6551.4e9) < 1e-5
    # errors are reported, not ignored
    assert lib.ma_write_yaml_report(C.byref(rep), str(tmp_path / "no" / "such" / "dir").encode(), None, 0) == -4
    assert lib.ma_write_yaml_report(None, None, None, 0) == -1


def test_driver_fails_loudly_without_a_gpu(lib, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert os.path.isfile(EXE)
    _write_inp(tmp_path / "miniaero.inp", dict(cases.REFERENCE_TESTS["3D_Sod_Serial"][0], ntimesteps=1))
    p = subprocess.run([EXE], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert p.returncode == 1 and "no CPU fallback" in p.stderr
    p = subprocess.run([EXE, "--input", "missing.inp"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert p.returncode == 1 and "missing.inp" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["3D_Sod_Serial", "FlatPlate_Serial"])
def test_driver_reproduces_the_reference_gold_file(lib, tmp_path, name):
    """miniaero.inp -> results.0 through the executable, against the reference's 6-digit gold file at the tolerance
    of its own test script (tests/<case>/<case>_test.sh), STRICT and FAST arithmetic; plus the YAML report."""
    import refrun
    inp, rel_tol, floor = cases.REFERENCE_TESTS[name]
    gold = parity.golden(name)["results_gold"]
    for arith in ("strict", "fast"):
        d = tmp_path / arith
        d.mkdir()
        _write_inp(d / "miniaero.inp", inp)
        p = subprocess.run([EXE, "--arith", arith], cwd=d, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stdout + p.stderr
        assert "Device Run time" in p.stdout and "Total elapsed time" in p.stdout and "Setup time" in p.stdout
        res = np.loadtxt(d / "results.0")
        assert res.shape == gold.shape
        assert refrun.numeric_text_diff(res, gold, rel_tol, floor) == 0
        assert len([f for f in os.listdir(d) if f.endswith(".yaml")]) == 1


@pytest.mark.gpu
def test_driver_without_output_uses_the_structured_constructor(lib, tmp_path):
    """output_results = 0: no host mesh is generated (ma_solver_create_structured); the run, its progress lines and the
    YAML report are the same."""
    inp = dict(cases.EXTRA["sod_o2_visc"], ntimesteps=3, output_results=0)
    _write_inp(tmp_path / "miniaero.inp", inp)
    p = subprocess.run([EXE], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "Device Run time" in p.stdout and "Setup time" in p.stdout and "4096 cells x 3 steps" in p.stdout
    assert not os.path.exists(tmp_path / "results.0")
    assert len([f for f in os.listdir(tmp_path) if f.endswith(".yaml")]) == 1


def _launch_ranks(args, n, cwd, env_extra=None, timeout=600):
    """n ranks of the driver started by this process (they share a parent, like the ranks of one torchrun agent)."""
    procs = []
    for r in range(n):
        env = dict(os.environ, WORLD_SIZE=str(n), RANK=str(r), LOCAL_RANK=str(r), MASTER_PORT="29517")
        env.pop("MINIAERO_RENDEZVOUS", None)
        env.update(env_extra or {})
        procs.append(subprocess.Popen([EXE] + args, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    return [(p.wait(timeout=timeout), *p.communicate()) for p in procs]


def test_rendezvous_survives_a_second_launch_and_a_stale_file(lib, tmp_path):
    """The NCCL-id rendezvous of the multi-rank driver (no GPU needed: --rendezvous-selftest exchanges a random token
    the same way).  Two launches in a row with the same MASTER_PORT must each agree on their OWN token, a stale file
    of an earlier aborted launch must not be read, and nothing may be left behind in /tmp."""
    assert os.path.isfile(EXE)
    stale = "/tmp/miniaero_rdv.29517.%d" % os.getpid()
    with open(stale, "wb") as f:
        f.write(b"\x55" * 128)          # what an aborted earlier launch of the same parent would leave
    tokens = []
    for _ in range(2):
        out = _launch_ranks(["--rendezvous-selftest"], 2, tmp_path, timeout=120)
        assert all(rc == 0 for rc, _, _ in out), out
        toks = {re.search(r"token ([0-9a-f]{16})", so).group(1) for _, so, _ in out}
        assert len(toks) == 1, out       # both ranks hold rank 0's token of THIS launch
        tokens.append(toks.pop())
    stale_hash = 1469598103934665603
    for b in b"\x55" * 128:
        stale_hash = ((stale_hash ^ b) * 1099511628211) % (1 << 64)
    assert tokens[0] != tokens[1] and "%016x" % stale_hash not in tokens
    assert not [f for f in os.listdir("/tmp") if f.startswith("miniaero_rdv.29517.%d" % os.getpid())]


@pytest.mark.gpu
def test_two_rank_driver_twice(lib, tmp_path):
    """Two launches of the 2-rank driver back to back (same port): each builds its own communicator and reproduces
    the reference's MPI build per rank (tests/golden/par_sod_o2_visc_2.npz) at the FAST tolerance."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    inp = dict(cases.PARALLEL["sod_o2_visc"][0], output_results=1)
    g = np.load(os.path.join(ROOT, "tests", "golden", "par_sod_o2_visc_2.npz"))
    for launch in range(2):
        d = tmp_path / ("launch%d" % launch)
        d.mkdir()
        _write_inp(d / "miniaero.inp", inp)
        out = _launch_ranks(["--precision", "17", "--no-yaml"], 2, d)
        assert all(rc == 0 for rc, _, _ in out), out
        for r in range(2):
            res = np.loadtxt(d / ("results.%d" % r))
            ref = g["r%d_step%d" % (r, inp["ntimesteps"])]
            linf, l2 = parity.field_errors(res[:, 3:], ref)
            assert linf <= parity.TOL_100_STEPS and l2 <= parity.TOL_100_STEPS
