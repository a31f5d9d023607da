"""TEST INFRASTRUCTURE — driver for the reference binaries in oracle/_ref/ (see build_ref.sh).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm may import this.
It runs the unmodified reference (built against the Kokkos stand-in) in a scratch directory holding
a `miniaero.inp`, and reads back (a) `results.<rank>` (6-digit text, what the reference's own tests
compare) and (b) the full-precision dumps written by the stand-in's deep_copy hook.
"""
import os
import re
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_MAGIC = 0x4F52454132303031


def ref_binary(kind="cell", omp=False):
    name = "miniAero.%s%s" % (kind, ".omp" if omp else "")
    path = os.path.join(REF_DIR, name)
    return path if os.path.isfile(path) and os.access(path, os.X_OK) else None


def write_inp(path, problem_type, lx, ly, lz, angle, nx, ny, nz, ntimesteps, dt, output_results,
              output_frequency, second_order, viscous):
    """The 9-line miniaero.inp format of Options.h:86-100."""
    with open(path, "w") as f:
        f.write("%d\n%r %r %r %r\n%d %d %d\n%d\n%r\n%d\n%d\n%d\n%d\n" % (
            problem_type, lx, ly, lz, angle, nx, ny, nz, ntimesteps, dt, output_results,
            output_frequency, second_order, viscous))


def read_dump(path):
    with open(path, "rb") as f:
        hdr = struct.unpack("<7q", f.read(56))
        assert hdr[0] == _MAGIC, path
        elem, rank, dims = hdr[1], hdr[2], hdr[3:3 + hdr[2]]
        dtype = {8: np.float64, 4: np.int32}[elem]
        data = np.frombuffer(f.read(), dtype=dtype, count=int(np.prod(dims)))
    return data.reshape(dims).copy()


def run_reference(inp, kind="cell", omp=False, threads=None, dump=True, workdir=None, timeout=3600):
    """Run the reference on the options dict `inp`; returns dict(stdout, run_time, results, dumps).

    dumps: list of (label, ndarray) in deep_copy order.  `inp` keys are write_inp's arguments.
    """
    exe = ref_binary(kind, omp)
    if exe is None:
        raise FileNotFoundError("oracle/_ref/miniAero.%s not built (run oracle/build_ref.sh where "
                                "/root/reference exists)" % kind)
    tmp = workdir or tempfile.mkdtemp(prefix="miniaero_ref_")
    write_inp(os.path.join(tmp, "miniaero.inp"), **inp)
    env = dict(os.environ)
    if dump:
        os.makedirs(os.path.join(tmp, "dump"), exist_ok=True)
        env["MINIAERO_DUMP_DIR"] = os.path.join(tmp, "dump")
    else:
        env.pop("MINIAERO_DUMP_DIR", None)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
        env.setdefault("OMP_PROC_BIND", "spread")
    p = subprocess.run([exe], cwd=tmp, env=env, capture_output=True, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError("reference failed: %s\n%s" % (p.returncode, p.stderr[-2000:]))
    out = {"stdout": p.stdout, "workdir": tmp}
    m = re.search(r"Device Run time:\s*([0-9.]+) seconds", p.stdout)
    out["run_time"] = float(m.group(1)) if m else None
    res = os.path.join(tmp, "results.0")
    out["results"] = np.loadtxt(res, ndmin=2) if os.path.isfile(res) else None
    dumps = []
    if dump:
        for name in sorted(os.listdir(env["MINIAERO_DUMP_DIR"])):
            dumps.append((name.split("_", 1)[1][:-4], read_dump(os.path.join(env["MINIAERO_DUMP_DIR"], name))))
    out["dumps"] = dumps
    return out


def run_reference_parallel(inp, nranks, dump=True, timeout=3600, labels="solution_n"):
    """Run the reference's WITH_MPI build (oracle/_ref/miniAero.cell.mpi, MPI = oracle/mpi_standin) as `nranks`
    cooperating processes.  Returns a list, per rank, of dict(results, dumps, stdout)."""
    exe = os.path.join(REF_DIR, "miniAero.cell.mpi")
    if not (os.path.isfile(exe) and os.access(exe, os.X_OK)):
        raise FileNotFoundError("oracle/_ref/miniAero.cell.mpi not built (run oracle/build_ref.sh)")
    tmp = tempfile.mkdtemp(prefix="miniaero_refmpi_")
    write_inp(os.path.join(tmp, "miniaero.inp"), **inp)
    os.makedirs(os.path.join(tmp, "mpi"))
    procs = []
    for r in range(nranks):
        env = dict(os.environ, MINIAERO_MPI_RANK=str(r), MINIAERO_MPI_SIZE=str(nranks),
                   MINIAERO_MPI_DIR=os.path.join(tmp, "mpi"), OMP_NUM_THREADS="1")
        env.pop("MINIAERO_DUMP_DIR", None)
        if dump:
            d = os.path.join(tmp, "dump%d" % r)
            os.makedirs(d)
            env["MINIAERO_DUMP_DIR"] = d
            env["MINIAERO_DUMP_LABELS"] = labels  # every ghost exchange deep_copies: keep only what is asked for
        procs.append(subprocess.Popen([exe], cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    out = []
    for r, p in enumerate(procs):
        so, se = p.communicate(timeout=timeout)
        if p.returncode != 0:
            for q in procs:
                if q.poll() is None:
                    q.kill()
            raise RuntimeError("reference rank %d failed: %s\n%s" % (r, p.returncode, se[-2000:]))
        res = os.path.join(tmp, "results.%d" % r)
        dumps = []
        if dump:
            d = os.path.join(tmp, "dump%d" % r)
            for name in sorted(os.listdir(d)):
                dumps.append((name.split("_", 1)[1][:-4], read_dump(os.path.join(d, name))))
        out.append({"stdout": so, "results": np.loadtxt(res, ndmin=2) if os.path.isfile(res) else None, "dumps": dumps,
                    "workdir": tmp})
    return out


def solution_from_dumps(dumps):
    """The conserved variables [ncells][5] copied to host at the end of Solve()
    (TimeSolverExplicitRK4.h:516, view label "solution_n")."""
    for label, arr in reversed(dumps):
        if label == "solution_n":
            return arr
    raise KeyError("solution_n not dumped (output_results must be non-zero)")


def numeric_text_diff(a, b, rel_tol=1e-3, floor=1e-6):
    """Python-3 restatement of tests/tools/numeric_text_diff:60-88 on parsed arrays: a field differs
    when either value exceeds `floor` and |a-b|/max(|a|,|b|) > rel_tol.  Returns the number of
    differing lines."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return max(a.shape[0], b.shape[0])
    big = (np.abs(a) > floor) | (np.abs(b) > floor)
    denom = np.maximum(np.abs(a), np.abs(b))
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(denom > 0, np.abs(a - b) / denom, 0.0)
    bad = big & (rel > rel_tol)
    return int(bad.any(axis=1).sum())


def unit_oracle(function, rows):
    """Evaluate one of the reference's device functions (oracle/unit_oracle.cpp: roe | viscous | primitives |
    venkat | vanalbada) on an [n, width] array of inputs; returns the [n, out_width] outputs."""
    exe = os.path.join(REF_DIR, "unit_oracle")
    if not os.path.isfile(exe):
        raise FileNotFoundError("%s not built (oracle/build_ref.sh needs /root/reference)" % exe)
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        rows.tofile(fin)
        subprocess.run([exe, function, fin, fout], check=True)
        out = np.fromfile(fout, dtype=np.float64)
    return out.reshape(rows.shape[0], -1)
