#!/bin/bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED miniAero reference sources, where they lie under
# /root/reference/kokkos, against the Kokkos stand-in in oracle/kokkos_standin/ (Kokkos itself is
# not vendored by the reference and not installed; see DESIGN.md "Oracle").  Outputs go to
# oracle/_ref/ only (git-ignored, but shipped to the GPU box by gpurun).  No reference source is
# copied into this repository.
#
#   miniAero.cell        -DCELL_FLUX      serial    : the parity oracle (deterministic slot-ordered gather)
#   miniAero.cell.omp    -DCELL_FLUX      -fopenmp  : CPU baseline ("reference source on an OpenMP loop")
#   miniAero.atomics     -DATOMICS_FLUX   serial    : the reference Makefile's default build (noise floor)
#   miniAero.atomics.omp -DATOMICS_FLUX   -fopenmp
#   miniAero.cell.mpi    -DCELL_FLUX -DWITH_MPI=1 over oracle/mpi_standin (N cooperating processes)
#   miniAero.cell.vanalbada  -DCELL_FLUX with the stencil limiter redirected to VanAlbadaLimiter (oracle/vanalbada_swap.h)
#   miniAero.b200        the reference's Main.C + mesh generator, solver class = include/reference_binding/TimeSolverB200.h
#                        (links libminiaero_b200.so: the compiled proof that the C ABI is a drop-in; GPU needed to run)
#   unit_oracle          oracle/unit_oracle.cpp: the reference's device functions (Roe, viscous, primitives, limiters)
#                        included from the reference headers and evaluated on arrays of inputs
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${MINIAERO_REFERENCE:-/root/reference/kokkos}"
OUT="$HERE/_ref"
if [ ! -f "$REF/Main.C" ]; then
  echo "build_ref.sh: reference sources not found at $REF (skipping; prebuilt $OUT is used if present)"
  exit 0
fi
mkdir -p "$OUT"
SRCS="Main.C Parallel3DMesh.C MeshProcessor.C Face.C Cell.C ElementTopo.C ElementTopoHexa8.C CopyGhost.C MemoryUsage.C"
# -ffp-contract=off: plain IEEE-754 evaluation in source order (no FMA contraction), the oracle's definition.
COMMON="-O3 -std=gnu++17 -w -ffp-contract=off -I$HERE/kokkos_standin -I$REF"
build() { # name, flags
  local name="$1"; shift
  if [ "$OUT/$name" -nt "$HERE/kokkos_standin/Kokkos_Core.hpp" ] && [ "$OUT/$name" -nt "$HERE/build_ref.sh" ] && [ "$OUT/$name" -nt "$HERE/mpi_standin/mpi.h" ] && [ "$OUT/$name" -nt "$HERE/vanalbada_swap.h" ]; then return; fi
  (cd "$REF" && g++ $COMMON "$@" $SRCS -o "$OUT/$name")
  echo "built $OUT/$name"
}
build miniAero.cell        -DCELL_FLUX
build miniAero.cell.omp    -DCELL_FLUX -fopenmp
build miniAero.atomics     -DATOMICS_FLUX
build miniAero.atomics.omp -DATOMICS_FLUX -fopenmp
# the reference's WITH_MPI path (block decomposition + ghost exchange) over the file-based MPI stand-in:
# full-precision per-rank results for the multi-GPU parity tests
build miniAero.cell.mpi    -DCELL_FLUX -DWITH_MPI=1 -I$HERE/mpi_standin
# the limiter the reference ships but never calls
# (only Main.C — the translation unit that instantiates the solver — sees the redirect; ElementTopoHexa8.C uses the
# host-side, non-template MathTools of MathTools.h, which must not meet MathToolsDevice.h)
VA="$OUT/miniAero.cell.vanalbada"
if [ ! "$VA" -nt "$HERE/vanalbada_swap.h" ] || [ ! "$VA" -nt "$HERE/kokkos_standin/Kokkos_Core.hpp" ] || [ ! "$VA" -nt "$HERE/build_ref.sh" ]; then
  TMPO="$(mktemp -d)"
  (cd "$REF" && g++ $COMMON -DCELL_FLUX -include "$HERE/vanalbada_swap.h" -c Main.C -o "$TMPO/Main.o" &&
   g++ $COMMON -DCELL_FLUX "$TMPO/Main.o" ${SRCS#Main.C } -o "$VA")
  rm -rf "$TMPO"
  echo "built $VA"
fi
# the reference's own Main.C / mesh generator with its solver class replaced by the B200 drop-in
# (include/reference_binding/: the binding a maintainer would add; tests/test_reference_binding.py runs it on the GPU)
ROOT="$(cd "$HERE/.." && pwd)"
B2="$OUT/miniAero.b200"
LIBDIR="$ROOT/miniaero_b200"
if [ -f "$LIBDIR/libminiaero_b200.so" ]; then
  if [ ! "$B2" -nt "$ROOT/include/reference_binding/TimeSolverB200.h" ] || [ ! "$B2" -nt "$ROOT/include/reference_binding/use_b200_solver.h" ] \
     || [ ! "$B2" -nt "$ROOT/include/miniaero_b200.h" ] || [ ! "$B2" -nt "$HERE/kokkos_standin/Kokkos_Core.hpp" ] || [ ! "$B2" -nt "$HERE/build_ref.sh" ]; then
    TMPO="$(mktemp -d)"
    (cd "$REF" && g++ $COMMON -DCELL_FLUX -I"$ROOT/include" -I"$ROOT/include/reference_binding" \
        -include "$ROOT/include/reference_binding/use_b200_solver.h" -c Main.C -o "$TMPO/Main.o" &&
     g++ $COMMON -DCELL_FLUX "$TMPO/Main.o" ${SRCS#Main.C } -o "$B2" -L"$LIBDIR" -lminiaero_b200 \
        -Wl,-rpath,'$ORIGIN/../../miniaero_b200')
    rm -rf "$TMPO"
    echo "built $B2"
  fi
else
  echo "build_ref.sh: $LIBDIR/libminiaero_b200.so not built yet: skipping miniAero.b200"
fi
# unit oracle: a driver of ours around the reference's headers (no reference source is copied)
if [ ! "$OUT/unit_oracle" -nt "$HERE/unit_oracle.cpp" ] || [ ! "$OUT/unit_oracle" -nt "$HERE/kokkos_standin/Kokkos_Core.hpp" ]; then
  g++ $COMMON -DCELL_FLUX "$HERE/unit_oracle.cpp" -o "$OUT/unit_oracle"
  echo "built $OUT/unit_oracle"
fi
