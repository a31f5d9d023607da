#!/usr/bin/env python
"""Benchmark of the explicit-RK4 finite-volume step (BASELINE.json metric: cell-updates/s, FP64, % HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one RK4 time step (4 stages) over the whole mesh.  One process per GPU; for N > 1 launch with
`python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N`.
torch / torch.distributed are plumbing only (process group, barrier, pinned host buffers); every number is
produced by libminiaero_b200.so through its C ABI.  There is no CPU fallback: without a B200 the GPU arm
fails.  `--impl reference` times the UNMODIFIED reference (oracle/_ref, its sources on an OpenMP loop) on
the host cores; it and the `cpu_baseline` leg are the only places this file executes anything of oracle/.

Workloads (all inputs are generated in code, data = "synthetic"):
  sod_o2_visc  3-D Sod tube, second order (Green-Gauss + Venkatakrishnan) + viscous flux, 512x512x256 cells
               PER GPU (BASELINE configs[3] restricted to N GPUs; at N = 1 it is configs[1]'s mesh with the
               viscous term on: the "full second-order viscous RK4 step" of the north star).  DEFAULT.
  sod_o2       BASELINE configs[1]: the same mesh, second order, inviscid.
  sod_o1       BASELINE configs[0]'s physics (first-order inviscid Roe, TimeSolverExplicitRK4.h:402-426
               <false, roe_flux, no_viscous_flux>) on the same mesh: one kernel per stage, 1 928 algorithmic bytes.
  flatplate    BASELINE configs[2]: viscous flat plate 1024x512x128 (NoSlip/Inflow/Tangent/Extrapolate).
  flatplate_strong  BASELINE configs[4]: the flat plate on a fixed 1024x512x512 global mesh split over the N GPUs
               ("scaling": "strong"; not part of the default run).
The default run also measures the other two single-GPU workloads for a few steps and reports them under
"also" (N = 1 only).
"""
import argparse
import json
import math
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ALGORITHMIC bytes (DESIGN.md "Roofline accounting"): doubles per cell per RK stage
DBL_SWEEP1 = 53            # gradient + min/max + limiter sweep (second order / viscous only)
DBL_SWEEP2_O2 = 91         # flux + gather + RK sweep, second order
DBL_SWEEP2_O1 = 59         # flux + gather + RK sweep, first order
BYTES_PER_CELL_UPDATE_O2 = 4 * 8 * (DBL_SWEEP1 + DBL_SWEEP2_O2) + 40   # 4648
BYTES_PER_CELL_UPDATE_O1 = 4 * 8 * DBL_SWEEP2_O1 + 40                  # 1928

WEAK_DIMS = {1: (512, 512, 256), 2: (1024, 512, 256), 4: (1024, 1024, 256), 8: (1024, 1024, 512)}


def workload_options(name, n_gpus, cells=None):
    """-> dict of miniaero.inp values (Options.h:91-99) for `name` on n_gpus GPUs (weak scaling)."""
    if name in ("sod_o2_visc", "sod_o2", "sod_o1"):
        g = cells or WEAK_DIMS[n_gpus]
        # the cell size of the 512x512x256 single-GPU mesh is kept as the mesh grows
        return dict(problem_type=0, lx=0.3048 * g[0] / 512.0, ly=1.0 * g[1] / 512.0, lz=1.0 * g[2] / 256.0, angle=0.0,
                    nx=g[0], ny=g[1], nz=g[2], dt=5e-7, second_order_space=0 if name == "sod_o1" else 1,
                    viscous=1 if name == "sod_o2_visc" else 0)
    if name == "flatplate":
        base = (1024, 512, 128)
        g = cells or tuple(b * s for b, s in zip(base, {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[n_gpus]))
        return dict(problem_type=1, lx=2.0 * g[0] / 1024.0, ly=0.032 * g[1] / 512.0, lz=1.0 * g[2] / 128.0, angle=0.0,
                    nx=g[0], ny=g[1], nz=g[2], dt=3e-8, second_order_space=1, viscous=1)
    if name == "flatplate_strong":
        # BASELINE configs[4]: the viscous flat plate on a FIXED 1024 x 512 x 512 global mesh (268 M cells) split over
        # the N GPUs (strong scaling; at N = 1 it needs ~162 GB of device and ~240 GB of host memory)
        g = cells or (1024, 512, 512)
        return dict(problem_type=1, lx=2.0 * g[0] / 1024.0, ly=0.032 * g[1] / 512.0, lz=1.0 * g[2] / 128.0, angle=0.0,
                    nx=g[0], ny=g[1], nz=g[2], dt=3e-8, second_order_space=1, viscous=1)
    raise SystemExit("unknown workload %r" % name)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------
# clocks
class ClockSampler:
    """nvidia-smi sampled every 100 ms during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        pw = [float(r[2]) for r in rows if re.match(r"^[0-9.]+$", r[2])]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(rows[0][1]) if rows else None,
                "power_w_max": max(pw) if pw else None, "samples": len(rows), "reasons": reasons}


# ---------------------------------------------------------------------------------------------------------
# the reference on the host cores (oracle/_ref; TEST INFRASTRUCTURE used here only as the timed baseline)
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def reference_sample_dims(opt, steps_total, budget_s, cores):
    """A bounded sample of the workload: the same physics on the largest mesh of a fixed ladder that the
    reference (~4e4 cell-updates/s/core second order, ~11 us/cell of mesh setup) finishes within budget_s."""
    ladder = [(32, 32, 16), (64, 32, 32), (64, 64, 32), (128, 64, 64), (128, 128, 64), (128, 128, 128), (256, 128, 128)]
    rate = 4.0e4 * cores * (1.0 if opt["second_order_space"] else 3.0)
    best = ladder[0]
    for d in ladder:
        c = d[0] * d[1] * d[2]
        if c * steps_total / rate + 12e-6 * c <= budget_s:
            best = d
    return best


def run_reference(workload, steps, warmup, budget_s, kind="cell", dims=None, threads=None):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refrun
    cores = threads or host_cores()
    base = workload_options(workload, 1)
    d = dims or reference_sample_dims(base, steps + warmup, budget_s, cores)
    full = (base["nx"], base["ny"], base["nz"])
    inp = dict(problem_type=base["problem_type"], lx=base["lx"] * d[0] / full[0], ly=base["ly"] * d[1] / full[1],
               lz=base["lz"] * d[2] / full[2], angle=base["angle"], nx=d[0], ny=d[1], nz=d[2], ntimesteps=steps + warmup,
               dt=base["dt"], output_results=0, output_frequency=10 ** 9, second_order=base["second_order_space"],
               viscous=base["viscous"])
    exe = refrun.ref_binary(kind, omp=True)
    if exe is None:
        return None
    tmp = tempfile.mkdtemp(prefix="miniaero_bench_ref_")
    log = os.path.join(tmp, "launch.log")
    os.environ["MINIAERO_LAUNCH_LOG"] = log
    try:
        refrun.run_reference(inp, kind=kind, omp=True, threads=cores, dump=False, workdir=tmp)
    finally:
        os.environ.pop("MINIAERO_LAUNCH_LOG", None)
    ends, functors = [], {}
    with open(log) as f:
        for line in f:
            p = line.split()
            if p[0] == "step_end":
                ends.append(float(p[1]))
            elif p[0] == "functor":
                functors[p[1]] = (int(p[2]), float(p[3]))
    # ends[0] = the copy before the time loop, ends[i] = end of time step i (TimeSolverExplicitRK4.h:335,488)
    assert len(ends) == steps + warmup + 1, (len(ends), steps, warmup)
    seconds = ends[steps + warmup] - ends[warmup]
    cells = d[0] * d[1] * d[2]
    top = sorted(functors.items(), key=lambda kv: -kv[1][1])[:4]
    total = sum(v[1] for v in functors.values()) or 1.0
    return {"value": cells * steps / seconds, "seconds": seconds, "cells": cells, "dims": list(d), "cores": cores,
            "exe": os.path.relpath(exe, ROOT),
            "top_functors": {re.sub(r"^\d+|IN6Kokkos.*$", "", k): round(v[1] / total, 3) for k, v in top}}


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = run_reference(args.workload, args.steps, args.warmup, budget_s=args.ref_budget)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/miniAero.cell.omp is not built "
                          "(oracle/build_ref.sh needs /root/reference; the prebuilt binary normally travels)"}))
        return 0
    opt = workload_options(args.workload, max(1, args.gpus))
    sample = ("%s physics on a %dx%dx%d mesh (%d cells), %d warm-up + %d timed RK4 steps, per-step times read from "
              "the Kokkos stand-in's launch log" % (args.workload, r["dims"][0], r["dims"][1], r["dims"][2], r["cells"],
                                                    args.warmup, args.steps))
    cpu_rec = {"value": r["value"], "unit": "cell-updates/s", "cores": r["cores"], "kind": "reference",
               "sample": sample, "cpu": cpu_model(), "top_functors": r["top_functors"]}
    # the same sample with the reference Makefile's default build (-DATOMICS_FLUX) and on ONE thread (a smaller
    # sample of the same physics): context for the all-cores figure above
    few = max(2, min(args.steps, 3))
    ra = run_reference(args.workload, few, 1, budget_s=args.ref_budget, kind="atomics", dims=tuple(r["dims"]))
    if ra is not None:
        cpu_rec["atomics_flux_build"] = {"value": ra["value"], "cores": ra["cores"], "steps": few,
                                         "what": "reference Makefile default -DATOMICS_FLUX (nondeterministic summation order)"}
    r1 = run_reference(args.workload, few, 1, budget_s=min(20.0, args.ref_budget), threads=1)
    if r1 is not None:
        cpu_rec["one_thread"] = {"value": r1["value"], "cores": 1, "steps": few, "dims": r1["dims"], "cells": r1["cells"]}
    line = {"impl": "reference", "metric": "cell-updates/sec (RK4 steps x cells) FP64", "value": r["value"],
            "unit": "cell-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * r["seconds"] / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            # the mesh this arm actually timed; `sample_of` names the GPU arm's workload it is a bounded sample of
            "config": {"workload": args.workload, "mesh": list(r["dims"]), "cells": r["cells"],
                       "sample_of": {"workload": args.workload, "mesh": [opt["nx"], opt["ny"], opt["nz"]],
                                     "cells_per_gpu": opt["nx"] * opt["ny"] * opt["nz"] // max(1, args.gpus),
                                     "why": "same physics, cell size and dt on a smaller mesh: the reference's ~5.5 KB per cell and "
                                            "its serial mesh set-up make the full size impractical on the host, so the two "
                                            "arms agree in everything but the number of cells"},
                       "second_order": opt["second_order_space"], "viscous": opt["viscous"],
                       "note": "reference = unmodified miniAero sources (-DCELL_FLUX, -O3 -fopenmp) on a Kokkos "
                               "stand-in whose parallel_for is an OpenMP static loop; CPU only"},
            "cpu_baseline": cpu_rec,
            "e2e": {"value": r["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0



# ---------------------------------------------------------------------------------------------------------
# parity of the very build that is about to be timed (tests/golden/*.npz = full-precision results of the unmodified
# reference; made by tests/golden/make_golden.py from oracle/_ref, re-derived by tests/test_oracle.py)
def parity_check(ma, comm, rank, world, local_rank, dist):
    """Small meshes, 1 / 2 / 100 RK4 steps, per rank against the reference's own results: its serial build at N = 1,
    its WITH_MPI build on the same block decomposition at N > 1 (CopyGhost.h:93-211, CopyGhost.C:41-79 are what the
    NCCL halo path replaces).  STRICT arithmetic must be bit-identical, FAST within the north-star tolerance on the
    global field's scale.  -> dict for the bench line; ok == False fails the run."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    import parity
    out = {"cases": {}, "tolerance": {"strict_ulp": 0, "fast_per_step": parity.TOL_PER_STEP, "fast_100_steps": parity.TOL_100_STEPS}}
    names = ["sod_o2_visc"] if world == 1 else [n for n in ("sod_o2_visc", "ramp_o2_visc", "3D_Sod_Parallel")
                                                if world in cases.PARALLEL[n][1]]
    worst_ulp, worst_fast, ok = 0, 0.0, True
    for name in names:
        if world == 1:
            inp, g = cases.EXTRA[name], parity.golden(name)
            key = lambda n: "cell_step%d" % n
        else:
            inp = cases.PARALLEL[name][0]
            g = np.load(os.path.join(parity.GOLDEN, "par_%s_%d.npz" % (name, world)))
            key = lambda n: "r%d_step%d" % (rank, n)
        entry = {}
        # FAST runs with shared cut faces forced ON: the configuration of the mesh that is timed (the default turns it
        # on from 2^20 cells); the small mesh is also run with it off and must give the same bits
        for arith, tag in ((ma.ARITH_STRICT, "strict"), (ma.ARITH_FAST, "fast")):
            for n, tol in ((1, parity.TOL_PER_STEP), (2, 2 * parity.TOL_PER_STEP), (100, parity.TOL_100_STEPS)):
                opt = ma.Options(**cases.opts_kwargs(dict(inp, ntimesteps=n)))
                solver = ma.TimeSolverExplicitRK4.from_options(opt, rank, world, device=local_rank, comm=comm, arith=arith,
                                                               share_cut_faces=1 if arith == ma.ARITH_FAST else 0)
                solver.initialize()
                solver.step(n)
                sol, ref = solver.solution(), g[key(n)]
                del solver
                if arith == ma.ARITH_FAST and n == 2:
                    plain = ma.TimeSolverExplicitRK4.from_options(opt, rank, world, device=local_rank, comm=comm, arith=arith,
                                                                  share_cut_faces=-1)
                    plain.initialize()
                    plain.step(n)
                    same = parity.max_ulp(plain.solution(), sol) == 0
                    del plain
                    if world > 1:
                        flags = [None] * world
                        dist.all_gather_object(flags, bool(same))
                        same = all(flags)
                    entry["shared_cut_faces_same_bits"] = bool(same)
                    ok = ok and same
                mine = {"ulp": parity.max_ulp(sol, ref), "parts": parity.field_error_parts(sol, ref)}
                rows = [mine]
                if world > 1:
                    rows = [None] * world
                    dist.all_gather_object(rows, mine)
                linf, l2 = parity.combine_parts([r["parts"] for r in rows])
                ulp = max(r["ulp"] for r in rows)
                if arith == ma.ARITH_STRICT:
                    entry["ulp_strict_step%d" % n] = ulp
                    worst_ulp = max(worst_ulp, ulp)
                    ok = ok and ulp == 0
                else:
                    entry["linf_fast_step%d" % n], entry["l2_fast_step%d" % n] = linf, l2
                    worst_fast = max(worst_fast, linf / tol, l2 / tol)
                    ok = ok and linf <= tol and l2 <= tol
        out["cases"][name] = entry
    out.update({"ulp_strict": worst_ulp, "fast_error_over_tolerance": worst_fast, "ok": bool(ok), "ranks": world,
                "reference": "tests/golden/%s (unmodified reference, %s)" % (
                    "<case>.npz" if world == 1 else "par_<case>_%d.npz" % world,
                    "-DCELL_FLUX serial build" if world == 1 else "-DCELL_FLUX -DWITH_MPI build, %d ranks" % world)})
    return out


# ---------------------------------------------------------------------------------------------------------
# host side of the end-to-end leg
def bind_to_gpu_numa_node(local_rank):
    """Run this rank — and therefore first-touch its pinned buffers — on the cores of its GPU's NUMA node.
    -> description for the bench line (a box that reports one node for every GPU has nothing to bind)."""
    info = {"numa_node": None, "bound": False}
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        info["numa_node"] = node
        nodes = [d for d in os.listdir("/sys/devices/system/node") if re.fullmatch(r"node\d+", d)]
        info["numa_nodes"] = len(nodes)
        if node >= 0 and len(nodes) > 1:
            cpus = set()
            for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["bound"], info["cpus"] = True, len(cpus)
    except Exception as ex:   # no sysfs, no such attribute: nothing to bind
        info["note"] = str(ex)[:120]
    return info


def host_link_bandwidth(torch, nbytes, barrier, max_over_ranks):
    """What the box can feed: every rank copies nbytes host->device and device->host between pinned memory and its GPU
    at the same time (two streams), all ranks together — the pattern of the end-to-end leg.  GB/s per rank and
    direction, from the slowest rank."""
    n = int(min(nbytes, 1 << 30)) // 8
    h_in = torch.empty(n, dtype=torch.float64, pin_memory=True).fill_(1.0)
    h_out = torch.empty(n, dtype=torch.float64, pin_memory=True)
    d_in = torch.empty(n, dtype=torch.float64, device="cuda")
    d_out = torch.ones(n, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ("h2d", "d2h", "both"):
        best = float("inf")
        for rep in range(3):
            barrier()
            t0 = time.perf_counter()
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
            s1.synchronize(); s2.synchronize()
            best = min(best, max_over_ranks(time.perf_counter() - t0))
        res[mode + "_gbs_per_rank"] = (2 if mode == "both" else 1) * n * 8 / best / 1e9
    res["bytes"] = n * 8
    return res


def source_hash():
    """Hash of the kernel sources: profiles/traffic.json is only quoted for the build it was measured on."""
    import hashlib
    h = hashlib.sha256()
    for f in ("kernels.cu", "physics.cuh", "kernels.h"):
        with open(os.path.join(ROOT, "miniaero_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


# ---------------------------------------------------------------------------------------------------------
def build_solver(ma, optd, rank, world, comm, local_rank, keep_mesh=False, mesh_path="structured", arith=0):
    """mesh_path "structured": ma_solver_create_structured (layout from (i, j, k), geometry evaluated on the device);
    "arrays": ma_mesh_generate + ma_solver_create (the reference-format arrays, re-packed and uploaded).  Same solver,
    bit for bit (tests/test_gpu_structured.py)."""
    opt = ma.Options(ntimesteps=1, output_results=0, output_frequency=10 ** 9, **optd)
    t0 = time.perf_counter()
    if mesh_path == "structured" and not keep_mesh:
        solver = ma.TimeSolverExplicitRK4.from_options(opt, rank, world, device=local_rank, comm=comm, arith=arith)
        t2 = time.perf_counter()
        n = [optd["nx"], optd["ny"], optd["nz"]]
        nproc, left = [1, 1, 1], world
        while left > 1:   # Parallel3DMesh.C:247-303: bisect the largest remaining dimension
            left //= 2
            d = max(range(3), key=lambda k: (n[k], -k))
            nproc[d] *= 2
            n[d] = [optd["nx"], optd["ny"], optd["nz"]][d] // nproc[d]
        info = {"mesh_seconds": 0.0, "layout_upload_seconds": t2 - t0, "owned_cells": solver.num_owned_cells,
                "ghost_cells": solver.num_ghosts, "nlocal": n, "nproc": nproc, "mesh_path": "structured"}
        return solver, opt, info
    mesh = ma.Parallel3DMesh.from_options(opt, rank, world).fillMeshData()
    t1 = time.perf_counter()
    solver = ma.TimeSolverExplicitRK4(mesh, opt, device=local_rank, comm=comm, arith=arith)
    t2 = time.perf_counter()
    info = {"mesh_seconds": t1 - t0, "layout_upload_seconds": t2 - t1, "owned_cells": mesh.num_owned_cells,
            "ghost_cells": mesh.num_ghosts, "nlocal": list(mesh.nlocal), "nproc": list(mesh.nproc), "mesh_path": "arrays"}
    if not keep_mesh:
        solver.release_mesh()
        del mesh
    return solver, opt, info


def time_steps(solver, steps, warmup, barrier):
    solver.initialize()
    solver.step(warmup)
    solver.synchronize()
    solver.reset_timing()
    solver.set_profiling(True)
    barrier()
    w0 = time.perf_counter()
    solver.step(steps)          # CUDA events on the solver's stream bracket exactly these K steps
    solver.synchronize()
    w1 = time.perf_counter()
    solver.set_profiling(False)
    t = solver.timing()
    t["wall_seconds"] = w1 - w0
    t["w0"], t["w1"] = w0, w1
    return t


def gpu_arm(args):
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the one-time host side of the solver construction (tile
    # topology, OpenMP) would then run single-threaded.  Give each rank its share of the cores — before libgomp loads.
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    if local_world > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1":
        os.environ["OMP_NUM_THREADS"] = str(max(1, host_cores() // local_world))
    import torch
    import torch.distributed as dist
    import miniaero_b200 as ma

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch N > 1 with torch.distributed.run)" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = ma.HaloComm.from_torch_distributed(local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    numa = bind_to_gpu_numa_node(local_rank)
    # ---- parity of this build on this box, before anything is timed (fails the run when it is not met)
    parity_rec = None
    if args.parity:
        parity_rec = parity_check(ma, comm, rank, world, local_rank, dist)
        if not parity_rec["ok"]:
            if rank == 0:
                sys.stderr.write("bench.py: PARITY FAILED, nothing is reported: %s\n" % json.dumps(parity_rec))
            if world > 1:
                dist.barrier()
                dist.destroy_process_group()
            return 3

    cells = tuple(args.cells) if args.cells else None
    optd = workload_options(args.workload, world, cells)
    # host memory gate: the host-side mesh + layout of a 64 M-cell block peaks near 0.9 KB/cell; ranks build in
    # groups that fit the free host memory
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    avail = -max_over_ranks(-float(avail))   # one figure for all ranks: the group size below must agree (barriers)
    per_rank = (320.0 if args.mesh_path == "structured" else 900.0) * optd["nx"] * optd["ny"] * optd["nz"] / world
    group = int(max(1, min(world, avail * 0.8 // max(per_rank, 1.0))))
    solver = None
    for g0 in range(0, world, group):
        if g0 <= rank < g0 + group:
            solver, opt, info = build_solver(ma, optd, rank, world, comm, local_rank, mesh_path=args.mesh_path)
        barrier()
    n_owned = info["owned_cells"]
    second = bool(optd["second_order_space"])

    sampler = ClockSampler(local_rank) if rank == 0 else None
    t = time_steps(solver, args.steps, args.warmup, barrier)
    clocks = sampler.stop(t["w0"], t["w1"]) if sampler else None
    step_s = max_over_ranks(t["step_seconds"])
    total_cells = sum_over_ranks(float(n_owned))
    value = total_cells * args.steps / step_s

    # ---- roofline of the dominant kernel (flux + gather + RK), CUDA events recorded over the timed region
    peak, peak_src = measured_peak()
    n_flux = 4 * args.steps
    flux_ms = 1e3 * t["flux_seconds"] / n_flux
    grad_ms = 1e3 * t["grad_seconds"] / n_flux
    dbl2 = DBL_SWEEP2_O2 if second else DBL_SWEEP2_O1
    flux_bytes = 8.0 * dbl2 * n_owned
    achieved = flux_bytes / (flux_ms * 1e-3) / 1e9
    traffic, traffic_note = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tp):
        with open(tp) as f:
            tj = json.load(f)
        key = "flux_rk_o2" if second else "flux_rk_o1"
        # dram__bytes_read.sum + dram__bytes_write.sum of one ncu capture (tools/measure_traffic.sh), per owned cell;
        # quoted only for the kernel sources it was captured on
        if key in tj and tj.get("source_hash") == source_hash():
            traffic = tj[key]["dram_bytes_per_cell"] * n_owned
            traffic_note = "ncu %s, %s" % (tj.get("captured", "?"), tj.get("recipe", "tools/measure_traffic.sh"))
        elif key in tj:
            traffic_note = "profiles/traffic.json was captured on other kernel sources (hash %s, now %s): not quoted" % (
                tj.get("source_hash"), source_hash())
    # FP64 pipe utilisation of the stage kernels: an ncu metric, quoted from the committed --set full capture of the
    # same kernel sources (profiles/pipes.json), never measured under the timed region
    fp64_pipe = None
    pp = os.path.join(ROOT, "profiles", "pipes.json")
    if second and optd["viscous"] and os.path.isfile(pp):
        with open(pp) as f:
            pj = json.load(f)
        if pj.get("source_hash") == source_hash():
            fp64_pipe = {"pct": pj.get("fp64_pipe_pct"), "source": pj.get("source")}
    bpcu = BYTES_PER_CELL_UPDATE_O2 if (second or optd["viscous"]) else BYTES_PER_CELL_UPDATE_O1
    roofline = {"bound": "hbm", "kernel": "flux_rk_tma_kernel (face fluxes + slot-ordered gather + RK stage update; launched once per pass of the shared-cut-face scheme, the time is the stage's launches together)", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": flux_bytes, "launch_ms": flux_ms, "fp64_pipe": fp64_pipe,
                "share_of_step": t["flux_seconds"] / t["step_seconds"],
                "flux_evaluations_per_cell": t["faces_evaluated"] / float(n_owned),
                "flux_evaluations_per_cell_unshared": t["tile_faces_total"] / float(n_owned),
                "grad_limiter_kernel": {"launch_ms": grad_ms, "algorithmic_bytes_per_launch": 8.0 * DBL_SWEEP1 * n_owned,
                                        "achieved": (8.0 * DBL_SWEEP1 * n_owned / (grad_ms * 1e-3) / 1e9) if grad_ms > 0 else None,
                                        "frac": (8.0 * DBL_SWEEP1 * n_owned / (grad_ms * 1e-3) / 1e9 / peak) if grad_ms > 0 else None},
                "whole_step": {"bytes_per_cell_update": bpcu, "achieved": bpcu * (value / world) / 1e9,
                               "frac": bpcu * (value / world) / 1e9 / peak, "frac_of_8TBs_nominal": bpcu * (value / world) / 8e12}}
    launches = int(t["kernel_launches"])
    halo = None
    if world > 1:   # what the exchange costs and how much of it the interior tiles hide (max over ranks; rank 0's counts)
        stages = 4 * args.steps
        halo = {"exchange_ms_per_stage": 1e3 * max_over_ranks(t["halo_seconds"]) / stages,
                "exposed_wait_ms_per_stage": 1e3 * max_over_ranks(t["halo_wait_seconds"]) / stages,
                "exposed_wait_share_of_step": max_over_ranks(t["halo_wait_seconds"]) / step_s,
                "interior_tiles": int(t["num_interior_tiles"]), "boundary_tiles": int(t["num_tiles"] - t["num_interior_tiles"]),
                "send_cells": int(t["num_send_cells"]), "recv_cells": int(t["num_recv_cells"]),
                "bytes_sent_per_stage": int(t["num_send_cells"]) * 25 * 8,
                "what": "two exchanges per stage on a second stream (stage state 5 doubles, gradient + limiter 20 doubles per "
                        "ghost): pack kernel, one ncclGroup of send/recv pairs, unpack kernel; exposed wait = time the "
                        "compute stream stalled before a boundary-tile launch (CUDA events)",
                "grad_launch_ms": grad_ms, "flux_launch_ms": flux_ms}

    # ---- end to end through the C ABI with HOST buffers: every step uploads the state from pinned host memory,
    # advances one RK4 step and reads the new state back (ma_solver_set_solution / _step / _get_solution)
    e2e, finite = None, True
    if args.e2e_steps > 0:   # (0: a diagnostic run without the end-to-end legs and their host / device buffers)
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        nbytes = n_owned * 5 * 8
        link = host_link_bandwidth(torch, nbytes, barrier, max_over_ranks)
        link["numa"] = numa
        link["concurrent_ranks"] = world
        hin = torch.empty(n_owned * 5, dtype=torch.float64, pin_memory=True)
        hout = torch.empty(n_owned * 5, dtype=torch.float64, pin_memory=True)
        solver.solution_into(hin.data_ptr())
        for _ in range(2):   # warm-up (allocates the device staging buffer)
            solver.set_solution(hin.data_ptr()); solver.step(1); solver.solution_into(hout.data_ptr())
        barrier()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            solver.set_solution(hin.data_ptr())
            solver.step(1)
            solver.solution_into(hout.data_ptr())   # synchronises the solver's stream
            hin, hout = hout, hin
        solver.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - w0)
        serial = {"value": total_cells * e2e_steps / e2e_s, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                  "what": "dependent chain, nothing overlapped: per step ma_solver_set_solution(pinned host) + "
                          "ma_solver_step(1) + ma_solver_get_solution(pinned host), each step's input = the previous output"}
        # The same three operations per step through ma_solver_submit: every step is an independent batch (its own
        # pinned input state, its own pinned output), so the upload of batch i+1 and the download of batch i-1 run on
        # their own streams while batch i is stepped.  Two input states (the solution at two different times) alternate.
        pipe_steps = max(args.steps, e2e_steps)
        try:
            for _ in range(2):   # warm-up (allocates the device staging buffers: four more copies of the state)
                solver.submit(hin.data_ptr(), hout.data_ptr(), 1)
            solver.synchronize()
            hin2 = hout.clone().pin_memory()
            houts = [hout, torch.empty_like(hout).pin_memory()]
            barrier()
            w0 = time.perf_counter()
            for i in range(pipe_steps):
                solver.submit((hin, hin2)[i & 1].data_ptr(), houts[i & 1].data_ptr(), 1)
            solver.synchronize()
            pipe_s = max_over_ranks(time.perf_counter() - w0)
            finite_pipe = bool(torch.isfinite(houts[0]).all().item() and torch.isfinite(houts[1]).all().item())
            e2e = {"value": total_cells * pipe_steps / pipe_s, "unit": "cell-updates/s", "h2d_bytes_per_step": nbytes,
                   "d2h_bytes_per_step": nbytes, "steps": pipe_steps, "ms_per_step": 1e3 * pipe_s / pipe_steps,
                   "what": "INDEPENDENT batches (ensemble members), one per step: ma_solver_submit(pinned host in, pinned host "
                           "out, 1) = upload of the batch's state, one RK4 step, download of the new state, consecutive batches "
                           "pipelined over three streams (fill and drain inside the timed region).  A time-stepping user who "
                           "moves the state across PCIe every step gets `serial_chain` (each step's input is the previous "
                           "output: nothing can overlap); one who calls Solve() for K steps, as the reference does, gets "
                           "`solve_call`",
                   "result_finite": finite_pipe, "serial_chain": serial}
            del hin2, houts
        except RuntimeError as ex:   # MiniAeroError is a RuntimeError; so is torch's failure to pin host memory
            if "memory" not in str(ex).lower() and "cudaMalloc" not in str(ex):
                raise
            # no room for the pipeline's staging buffers next to a solver that fills the device (every rank sizes alike):
            # the dependent chain is the end-to-end number then
            e2e = {"value": serial["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": nbytes,
                   "d2h_bytes_per_step": nbytes, "steps": serial["steps"], "ms_per_step": serial["ms_per_step"],
                   "what": serial["what"], "pipelined_unavailable": str(ex)[:200], "serial_chain": serial}
        # the reference's own call shape (Main.C:139-141): one Solve() of K steps, state up once, result down once
        barrier()
        w0 = time.perf_counter()
        solver.set_solution(hin.data_ptr()); solver.step(args.steps); solver.solution_into(hout.data_ptr())
        solver.synchronize()
        solve_s = max_over_ranks(time.perf_counter() - w0)
        e2e["solve_call"] = {"value": total_cells * args.steps / solve_s, "steps": args.steps,
                             "what": "one upload + K steps + one download (the reference's Solve() shape)"}
        # what the host can feed: the per-step traffic of this leg at the copy rates measured above (all ranks at once)
        both = link["both_gbs_per_rank"] / 2.0   # per direction while both directions run
        e2e["host_link"] = link
        e2e["host_link"]["ceiling"] = {
            "pipelined": total_cells / max(step_s / args.steps, nbytes / (both * 1e9)),
            "serial_chain": total_cells / (step_s / args.steps + nbytes / (link["h2d_gbs_per_rank"] * 1e9)
                                           + nbytes / (link["d2h_gbs_per_rank"] * 1e9)),
            "what": "cell-updates/s if every rank's upload and download ran at the measured pinned-copy rates: pipelined = "
                    "max(device step, transfer at the bidirectional rate); serial = device step + upload + download"}
        finite = bool(torch.isfinite(hout).all().item())
        del hin, hout

    # ---- the other single-GPU workloads, a few steps each (N = 1, default workload only)
    also = {}
    if world == 1 and args.also and not args.cells:
        del solver
        torch.cuda.empty_cache()
        for name in ("sod_o2", "flatplate", "sod_o1"):
            if name == args.workload:
                continue
            s2, _, i2 = build_solver(ma, workload_options(name, 1), 0, 1, None, local_rank, mesh_path=args.mesh_path)
            t2 = time_steps(s2, max(3, args.steps // 4), 3, barrier)
            v2 = i2["owned_cells"] * t2["steps"] / t2["step_seconds"]
            b2 = BYTES_PER_CELL_UPDATE_O1 if name == "sod_o1" else BYTES_PER_CELL_UPDATE_O2
            also[name] = {"value": v2, "unit": "cell-updates/s", "steps": int(t2["steps"]), "cells": i2["owned_cells"],
                          "ms_per_step": 1e3 * t2["step_seconds"] / t2["steps"],
                          "bytes_per_cell_update": b2, "whole_step_roofline_frac": b2 * v2 / 1e9 / peak,
                          "flux_launch_ms": 1e3 * t2["flux_seconds"] / (4 * t2["steps"])}
            del s2
        # the bit-for-bit mode (MA_ARITH_STRICT: IEEE evaluation in the reference's order, no FMA, gather kernels), the
        # default workload on a quarter-size mesh: what exact reproduction of the reference's -DCELL_FLUX build costs
        if args.workload == "sod_o2_visc":
            o3 = workload_options("sod_o2_visc", 1, (512, 256, 128))
            s3, _, i3 = build_solver(ma, o3, 0, 1, None, local_rank, mesh_path=args.mesh_path, arith=ma.ARITH_STRICT)
            t3 = time_steps(s3, 3, 3, barrier)
            v3 = i3["owned_cells"] * t3["steps"] / t3["step_seconds"]
            also["sod_o2_visc_strict_arith"] = {"value": v3, "unit": "cell-updates/s", "steps": int(t3["steps"]),
                                                "cells": i3["owned_cells"],
                                                "ms_per_step": 1e3 * t3["step_seconds"] / t3["steps"],
                                                "what": "MA_ARITH_STRICT: bit for bit the reference's -DCELL_FLUX results"}
            del s3

    # ---- strong scaling (BASELINE configs[4]): the viscous flat plate on a FIXED 1024 x 512 x 512 mesh over the N GPUs
    strong = None
    if args.strong and args.workload != "flatplate_strong" and not args.cells:
        try:
            del solver
        except NameError:
            pass
        torch.cuda.empty_cache()
        so = workload_options("flatplate_strong", world)
        need = 600.0 * so["nx"] * so["ny"] * so["nz"] / world   # ~575 B/cell of device memory
        free_dev = -max_over_ranks(-float(torch.cuda.mem_get_info()[0]))   # the least over the ranks: one decision for all
        if need > 0.95 * free_dev:
            strong = {"skipped": "%.0f GB per GPU needed, %.0f GB free" % (need / 1e9, free_dev / 1e9)}
        else:
            per_rank4 = (320.0 if args.mesh_path == "structured" else 900.0) * so["nx"] * so["ny"] * so["nz"] / world
            group4 = int(max(1, min(world, avail * 0.8 // max(per_rank4, 1.0))))
            s4, err4 = None, None
            for g0 in range(0, world, group4):
                if g0 <= rank < g0 + group4:
                    try:
                        s4, _, i4 = build_solver(ma, so, rank, world, comm, local_rank, mesh_path=args.mesh_path)
                    except RuntimeError as ex:   # out of memory on this rank: the leg is dropped on every rank, not the line
                        err4 = str(ex)[:200]
                barrier()
            if max_over_ranks(1.0 if err4 else 0.0) > 0.0:
                strong = {"failed": err4 or "another rank could not build its block"}
                s4 = None
                torch.cuda.empty_cache()
        if strong is None:
            t4 = time_steps(s4, max(3, args.steps // 2), 3, barrier)
            st4 = max_over_ranks(t4["step_seconds"])
            cells4 = so["nx"] * so["ny"] * so["nz"]
            strong = {"workload": "flatplate_strong", "mesh": [so["nx"], so["ny"], so["nz"]], "cells": cells4,
                      "value": cells4 * t4["steps"] / st4, "unit": "cell-updates/s", "steps": int(t4["steps"]),
                      "ms_per_step": 1e3 * st4 / t4["steps"], "cells_per_gpu": i4["owned_cells"], "blocks": i4["nproc"],
                      "setup_seconds": round(i4["layout_upload_seconds"], 2),
                      "halo_exposed_wait_ms_per_step": 1e3 * max_over_ranks(t4["halo_wait_seconds"]) / t4["steps"]}
            ref_file = os.path.join(ROOT, "profiles", "strong_scaling.json")
            if os.path.isfile(ref_file):   # efficiency against the smallest GPU count that holds the mesh (committed line)
                with open(ref_file) as f:
                    base = json.load(f).get("base")
                if base and base.get("value"):
                    strong["efficiency"] = strong["value"] / (base["value"] * world / base["n_gpus"])
                    strong["efficiency_vs"] = base
            del s4

    cpu = None
    if rank == 0 and world == 1 and args.cpu_baseline:
        r = run_reference(args.workload, 3, 1, budget_s=args.cpu_budget)
        if r is not None:
            cpu = {"value": r["value"], "unit": "cell-updates/s", "cores": r["cores"], "kind": "reference",
                   "cpu": cpu_model(), "top_functors": r["top_functors"],
                   "sample": "%s physics on a %dx%dx%d mesh (%d cells), 1 warm-up + 3 timed RK4 steps of the unmodified "
                             "reference (-DCELL_FLUX, OpenMP static loop over %d threads)" % (
                                 args.workload, r["dims"][0], r["dims"][1], r["dims"][2], r["cells"], r["cores"])}
            # the reference Makefile's default build (-DATOMICS_FLUX: atomic face-to-cell accumulation), same sample
            ra = run_reference(args.workload, 3, 1, budget_s=args.cpu_budget, kind="atomics", dims=tuple(r["dims"]))
            if ra is not None:
                cpu["atomics_flux_build"] = {"value": ra["value"], "what": "the same sample with the reference "
                                             "Makefile's default -DATOMICS_FLUX (nondeterministic summation order)"}

    if rank == 0:
        line = {"metric": "cell-updates/sec (RK4 steps x cells) FP64", "value": value, "unit": "cell-updates/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * step_s / args.steps,
                "higher_is_better": True, "scaling": "strong" if args.workload.endswith("_strong") else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": args.workload, "mesh": [optd["nx"], optd["ny"], optd["nz"]],
                           "cells_per_gpu": n_owned, "ghost_cells_rank0": info["ghost_cells"], "blocks": info["nproc"],
                           "second_order": optd["second_order_space"], "viscous": optd["viscous"], "dt": optd["dt"],
                           "arith": "fast (FMA contraction, reciprocal multiplication); parity vs the reference in tests/",
                           "l2": "no flush needed: %.1f GB of solver state per GPU >> 126 MB L2" % (t["device_bytes"] / 1e9),
                           "device_bytes": int(t["device_bytes"]), "tiles": int(t["num_tiles"]),
                           "setup_seconds": {k: round(v, 2) for k, v in info.items() if k.endswith("_seconds")},
                           "mesh_path": info["mesh_path"]},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "wall_ms_per_step": 1e3 * t["wall_seconds"] / args.steps, "result_finite": finite,
                "parity": parity_rec}
        if world > 1:
            line["halo"] = halo
        if strong is not None:
            line["strong"] = strong
        if also:
            line["also"] = also
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sod_o2_visc", choices=["sod_o2_visc", "sod_o2", "sod_o1", "flatplate", "flatplate_strong"])
    ap.add_argument("--cells", type=int, nargs=3, default=None, help="override the GLOBAL mesh (debugging only)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--mesh-path", default="structured", choices=["structured", "arrays"],
                    help="how the solver is constructed (set-up only; the timed steps are the same kernels)")
    ap.add_argument("--no-also", dest="also", action="store_false")
    ap.add_argument("--no-parity", dest="parity", action="store_false", help="skip the parity check of this build")
    ap.add_argument("--strong", dest="strong", action="store_true", default=None,
                    help="also time the fixed-size 1024x512x512 flat plate (default: on; skipped when it does not fit)")
    ap.add_argument("--no-strong", dest="strong", action="store_false")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the cpu_baseline leg")
    ap.add_argument("--ref-budget", type=float, default=90.0, help="seconds of CPU work for --impl reference")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.strong is None:
        args.strong = True
    # stdout carries exactly one JSON line: everything else a library prints there (NCCL's version banner, ...)
    # goes to stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    return reference_arm(args) if args.impl == "reference" else gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
