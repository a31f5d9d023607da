"""In-tree build of the C-ABI shared library (nvcc for sm_100a; cross-compiles without a GPU).

    python -m miniaero_b200.build [--force] [--verbose]

Outputs (git-ignored, shipped to the GPU box by gpurun):
    miniaero_b200/libminiaero_b200.so     the C-ABI library (include/miniaero_b200.h)
    miniaero_b200/miniaero                host driver executable (the reference's Main.C shape)
    miniaero_b200/build/ptxas.log         registers / spills / shared memory per kernel (-Xptxas -v)
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libminiaero_b200.so")
EXE = os.path.join(HERE, "miniaero")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off,-fopenmp",
               "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]
GXX_COMMON = ["-O3", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-Wall",
              "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _cuda_include():
    return os.path.join(os.path.dirname(os.path.dirname(_nvcc())), "include")


def _newer(target, sources):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, log=None, verbose=False):
    if verbose:
        print(" ".join(cmd), flush=True)
    p = subprocess.run(cmd, capture_output=True, text=True)
    if log is not None:
        log.write("$ " + " ".join(cmd) + "\n" + p.stdout + p.stderr + "\n")
    if p.returncode != 0:
        sys.stderr.write(p.stdout + p.stderr)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    return p


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(ROOT, "include", "miniaero_b200.h"))
    headers.append(os.path.abspath(__file__))
    objs = []
    with open(os.path.join(BUILD, "ptxas.log"), "a") as log:
        # the physics kernels, twice: FMA-contracted (fast) and strict IEEE order (strict)
        for tag, extra in (("fast", ["-fmad=true"]), ("strict", ["-fmad=false", "-DMA_STRICT=1"])):
            src = os.path.join(CSRC, "kernels.cu")
            obj = os.path.join(BUILD, "kernels_%s.o" % tag)
            if force or _newer(obj, [src] + headers):
                _run([nvcc] + ARCH + NVCC_COMMON + ["-Xptxas", "-v"] + extra + ["-c", src, "-o", obj], log, verbose)
            objs.append(obj)
        # device-side mesh geometry: the host generator's expressions, no FMA contraction (same bits as the host)
        src = os.path.join(CSRC, "geom_kernels.cu")
        obj = os.path.join(BUILD, "geom_kernels.o")
        if force or _newer(obj, [src] + headers):
            _run([nvcc] + ARCH + NVCC_COMMON + ["-Xptxas", "-v", "-fmad=false", "-c", src, "-o", obj], log, verbose)
        objs.append(obj)
        for name in ("solver.cu", "topology_kernels.cu"):
            src = os.path.join(CSRC, name)
            obj = os.path.join(BUILD, name.replace(".cu", ".o"))
            if force or _newer(obj, [src] + headers):
                _run([nvcc] + ARCH + NVCC_COMMON + ["-Xptxas", "-v", "-c", src, "-o", obj], log, verbose)
            objs.append(obj)
        for name in ("host_common.cpp", "host_mesh.cpp", "layout.cpp", "comm.cpp", "host_report.cpp"):
            src = os.path.join(CSRC, name)
            if not os.path.isfile(src):
                continue
            obj = os.path.join(BUILD, name.replace(".cpp", ".o"))
            if force or _newer(obj, [src] + headers):
                _run(["g++"] + GXX_COMMON + ["-I" + _cuda_include(), "-c", src, "-o", obj], log, verbose)
            objs.append(obj)
        if force or _newer(LIB, objs):
            _run([nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xcompiler", "-fopenmp", "-lgomp", "-ldl"], log,
                 verbose)
        main_src = os.path.join(CSRC, "main.cpp")
        if os.path.isfile(main_src) and (force or _newer(EXE, [main_src, LIB] + headers)):
            _run(["g++"] + GXX_COMMON + [main_src, "-o", EXE, "-L" + HERE, "-lminiaero_b200",
                                         "-Wl,-rpath,$ORIGIN"], log, verbose)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(LIB)
