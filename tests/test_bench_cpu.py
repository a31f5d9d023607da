"""bench.py without a GPU: the reference arm's JSON line (the contract's keys), the workloads against BASELINE.json's
configs, and the loud failure of the GPU arm."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-budget", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "cell-updates/s" and d["dtype"] == "f64" and d["value"] > 0
    assert d["vs_baseline"] is None and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "sod_o2_visc"
    # the arm says which mesh it timed and which workload that is a bounded sample of
    assert d["config"]["cells"] == d["config"]["mesh"][0] * d["config"]["mesh"][1] * d["config"]["mesh"][2]
    assert d["config"]["sample_of"]["mesh"] == [512, 512, 256] and d["config"]["sample_of"]["cells_per_gpu"] == 67108864
    assert d["cpu_baseline"]["one_thread"]["cores"] == 1 and d["cpu_baseline"]["atomics_flux_build"]["value"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_workloads_are_the_baseline_configs():
    import bench
    o = bench.workload_options("sod_o2", 1)
    assert (o["nx"], o["ny"], o["nz"]) == (512, 512, 256) and o["second_order_space"] == 1 and o["viscous"] == 0
    o = bench.workload_options("flatplate", 1)
    assert (o["nx"], o["ny"], o["nz"]) == (1024, 512, 128) and o["problem_type"] == 1 and o["viscous"] == 1
    for n, dims in bench.WEAK_DIMS.items():   # configs[3]: ~64 M cells per GPU under the reference's block arrangement
        o = bench.workload_options("sod_o2_visc", n)
        assert (o["nx"], o["ny"], o["nz"]) == dims and o["nx"] * o["ny"] * o["nz"] == n * 67108864
        assert abs(o["lx"] / o["nx"] - 0.3048 / 512) < 1e-15   # the cell size is kept as the mesh grows
    o = bench.workload_options("flatplate_strong", 8)          # configs[4]: fixed 268 M cells
    assert o["nx"] * o["ny"] * o["nz"] == 268435456
    assert bench.BYTES_PER_CELL_UPDATE_O2 == 4648 and bench.BYTES_PER_CELL_UPDATE_O1 == 1928
    o = bench.workload_options("sod_o1", 1)                    # configs[0]'s physics at the benchmark size
    assert (o["nx"], o["ny"], o["nz"]) == (512, 512, 256) and o["second_order_space"] == 0 and o["viscous"] == 0


def test_traffic_is_quoted_only_for_the_sources_it_was_measured_on():
    """profiles/traffic.json carries the hash of the kernel sources of its ncu capture; bench.py drops the figure
    (roofline.traffic = null with a note) when the sources have moved on."""
    import bench
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
        tj = json.load(f)
    assert len(bench.source_hash()) == 16 and "source_hash" in tj and tj["flux_rk_o2"]["dram_bytes_per_cell"] > 0
    # the committed captures belong to the committed kernels: both figures are quoted by the next bench run
    assert tj["source_hash"] == bench.source_hash()
    with open(os.path.join(ROOT, "profiles", "pipes.json")) as f:
        pj = json.load(f)
    assert pj["source_hash"] == bench.source_hash() and all(0 < v < 100 for v in pj["fp64_pipe_pct"].values())


def test_gpu_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert p.returncode != 0 and "no CPU fallback" in (p.stdout + p.stderr)
