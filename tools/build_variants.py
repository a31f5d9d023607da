"""Build experiment variants of the C-ABI library: the FAST kernels recompiled with different build-time knobs
(kernels.cu: MA_C128_FT, MA_C128_FB, MA_FLUX_RK_STAGED, MA_FLUX_XC, MA_FLUX_PREFETCH_AHEAD, ...), linked against the
regular objects.  Output: miniaero_b200/variants/libminiaero_b200_<tag>.so (git-ignored; travels with gpurun).
Select one at run time with MINIAERO_B200_LIB=<path> (developer knob of miniaero_b200/_abi.py).

    python tools/build_variants.py            # all variants below
    python tools/build_variants.py tag ...    # some
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miniaero_b200 import build as B  # noqa: E402

OLD = ["-DMA_C128_FB=3", "-DMA_FLUX_RK_STAGED=1", "-DMA_FLUX_XC=1"]   # the configuration before round 1e
VARIANTS = {
    # tag: extra -D flags (defaults: 128 threads x 4 CTAs per SM, RK operands read directly, no staged second record)
    "default": [],
    "pipe": [],                                    # + MINIAERO_FLUX_KERNEL=pipe at run time (tools/experiments/flux_pipe.cuh)
    "pg1": ["-DMA_PIPE_GATHER=1"],
    "t128b3": OLD,
    "t160b3": ["-DMA_C128_FT=160"] + OLD,
    "t192b3": ["-DMA_C128_FT=192"] + OLD,
    "t160b4": ["-DMA_C128_FT=160", "-DMA_C128_FB=4"],
    "t192b3n": ["-DMA_C128_FT=192", "-DMA_C128_FB=3"],       # RK operands direct (unlike t192b3)
    "t192b4": ["-DMA_C128_FT=192", "-DMA_C128_FB=4"],
    "t256b3": ["-DMA_C128_FT=256", "-DMA_C128_FB=3"],
    "t256b4": ["-DMA_C128_FT=256", "-DMA_C128_FB=4"],
    "t128b4p": ["-DMA_FLUX_PREFETCH_AHEAD=592"],
    "p296": ["-DMA_FLUX_PREFETCH_AHEAD=296", "-DMA_HEADER_AHEAD=1024"],
    "p592h": ["-DMA_FLUX_PREFETCH_AHEAD=592", "-DMA_HEADER_AHEAD=1536"],
    "p592g": ["-DMA_FLUX_PREFETCH_AHEAD=592", "-DMA_HEADER_AHEAD=1536", "-DMA_FLUX_PREFETCH_MIN_BYTES=2048"],
    "h1024": ["-DMA_HEADER_AHEAD=1024"],
    "c256b2": ["-DMA_C256_FB=2"],
    "t128b3p": ["-DMA_FLUX_PREFETCH_AHEAD=444"] + OLD,
    "g5": ["-DMA_C128_GB1=5"],
    "gsc0": ["-DMA_GRAD_STAGE_CELL=0"],
    "gsc1": ["-DMA_GRAD_STAGE_CELL=1"],   # gradient kernel: slot maps, volume, centroid arrive with the bulk copies
    "xg_copyonly": ["-DMA_GRAD_EXPERIMENT=1"],
    "xg_computeonly": ["-DMA_GRAD_EXPERIMENT=2"],
    "x_copyonly": ["-DMA_FLUX_EXPERIMENT=1"],
    "x_computeonly": ["-DMA_FLUX_EXPERIMENT=2"],
    "x_merged": ["-DMA_FLUX_EXPERIMENT=3"] + OLD,
}


EXPERIMENT_TAGS = ("x", "pg", "pipe")   # variants that need tools/experiments/*.patch applied to a copy of csrc/


def experiment_sources():
    """A scratch copy of csrc/ with the timing-experiment branches (MA_FLUX_EXPERIMENT / MA_GRAD_EXPERIMENT) and the
    pipelined flux kernel (MINIAERO_FLUX_KERNEL=pipe) patched back in: they are kept out of the product sources."""
    import shutil
    dst = os.path.join(B.BUILD, "experiment_src")
    shutil.rmtree(dst, ignore_errors=True)
    shutil.copytree(B.CSRC, dst)
    exp = os.path.join(ROOT, "tools", "experiments")
    shutil.copy(os.path.join(exp, "flux_pipe.cuh"), dst)
    for patch, target in (("timing_and_pipe_experiments.patch", "kernels.cu"), ("solver_pipe_variant.patch", "solver.cu")):
        p = subprocess.run(["patch", "-s", os.path.join(dst, target), os.path.join(exp, patch)], capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("tools/experiments/%s no longer applies to csrc/%s (regenerate it): %s" % (patch, target, p.stdout + p.stderr))
    return dst


def build_one(tag):
    out_dir = os.path.join(B.HERE, "variants")
    os.makedirs(out_dir, exist_ok=True)
    obj = os.path.join(B.BUILD, "kernels_fast_%s.o" % tag)
    lib = os.path.join(out_dir, "libminiaero_b200_%s.so" % tag)
    nvcc = B._nvcc()
    csrc = B.CSRC
    extra_objs = {}
    if tag.startswith(EXPERIMENT_TAGS):
        csrc = EXP_SRC[0]
        sobj = os.path.join(B.BUILD, "solver_exp.o")
        common = [a.replace(B.CSRC, csrc) for a in B.NVCC_COMMON]
        p = subprocess.run([nvcc] + B.ARCH + common + ["-c", os.path.join(csrc, "solver.cu"), "-o", sobj], capture_output=True, text=True)
        if p.returncode != 0:
            return tag, p.stderr[-2000:]
        extra_objs["solver.o"] = sobj
    common = [a.replace(B.CSRC, csrc) for a in B.NVCC_COMMON]
    p = subprocess.run([nvcc] + B.ARCH + common + ["-Xptxas", "-v", "-fmad=true"] + VARIANTS[tag] +
                       ["-c", os.path.join(csrc, "kernels.cu"), "-o", obj], capture_output=True, text=True)
    if p.returncode != 0:
        return tag, p.stderr[-2000:]
    info = []
    lines = p.stderr.splitlines()
    for i, ln in enumerate(lines):
        if "Function properties" in ln and "flux_rk_tma_kernelILb1ELb1ENS_7TileCapILi128" in ln:
            info = [lines[i + 1].strip(), lines[i + 2].strip()]
    objs = [obj] + [extra_objs.get(n, os.path.join(B.BUILD, n)) for n in ("kernels_strict.o", "geom_kernels.o", "solver.o", "topology_kernels.o",
                                                                        "host_common.o", "host_mesh.o", "layout.o", "comm.o",
                                                                        "host_report.o")]
    p = subprocess.run([nvcc] + B.ARCH + ["-shared", "-o", lib] + objs + ["-Xcompiler", "-fopenmp", "-lgomp", "-ldl"],
                       capture_output=True, text=True)
    if p.returncode != 0:
        return tag, p.stderr[-2000:]
    return tag, " | ".join(info)


EXP_SRC = [None]

if __name__ == "__main__":
    B.build()
    tags = sys.argv[1:] or list(VARIANTS)
    if any(t.startswith(EXPERIMENT_TAGS) for t in tags):
        EXP_SRC[0] = experiment_sources()
    with ThreadPoolExecutor(max_workers=4) as ex:
        for tag, msg in ex.map(build_one, tags):
            print(tag, "::", msg, flush=True)
