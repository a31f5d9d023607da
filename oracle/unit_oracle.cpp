// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Unit oracle: evaluates the reference's own device functions — included from where they lie under
// /root/reference/kokkos, compiled against oracle/kokkos_standin with -ffp-contract=off — on arrays of inputs.
//   roe        roe_flux<Device>::compute_flux               Roe_Flux.h:49-265
//   viscous    newtonian_viscous_flux<Device>::compute_flux Viscous_Flux.h:65-98
//   primitives ComputePrimitives<Device>                    GasModel.h:70-90
//   venkat     VenkatLimiter<Device>::limit                 VenkatLimiter.h:45-73
//   vanalbada  VanAlbadaLimiter<Device>::limit              VanAlbadaLimiter.h:45-65
// Usage: unit_oracle <function> <in.bin> <out.bin>; in.bin = n rows of doubles (row layouts below), out.bin = n rows.
// Built by oracle/build_ref.sh into oracle/_ref/unit_oracle; used by tests/golden/make_golden_unit.py and tests/.
#include <Kokkos_Core.hpp>

#include <cstdio>
#include <cstring>
#include <vector>

#include "GasModel.h"
#include "Roe_Flux.h"
#include "VanAlbadaLimiter.h"
#include "VenkatLimiter.h"
#include "Viscous_Flux.h"

typedef Kokkos::DefaultExecutionSpace Device;

int main(int argc, char **argv) {
  if (argc != 4) return fprintf(stderr, "usage: unit_oracle roe|viscous|primitives|venkat|vanalbada in.bin out.bin\n"), 2;
  const std::string fn = argv[1];
  int in_w, out_w;
  if (fn == "roe") in_w = 19, out_w = 5;           // Vl[5] Vr[5] normal[3] tangent[3] binormal[3] -> flux[5]
  else if (fn == "viscous") in_w = 23, out_w = 5;  // grad[5][3] V[5] a[3] -> vflux[5]
  else if (fn == "primitives") in_w = 5, out_w = 5;
  else if (fn == "venkat") in_w = 4, out_w = 1;    // dumax dumin du deltax3
  else if (fn == "vanalbada") in_w = 3, out_w = 1; // dumax dumin du
  else return fprintf(stderr, "unknown function %s\n", argv[1]), 2;
  FILE *f = fopen(argv[2], "rb");
  if (!f) return perror(argv[2]), 1;
  fseek(f, 0, SEEK_END);
  const long bytes = ftell(f);
  fseek(f, 0, SEEK_SET);
  const long n = bytes / (8 * in_w);
  std::vector<double> in((size_t)n * in_w), out((size_t)n * out_w);
  if (fread(in.data(), 8, in.size(), f) != in.size()) return fprintf(stderr, "short read\n"), 1;
  fclose(f);
  roe_flux<Device> roe;
  newtonian_viscous_flux<Device> visc;
  for (long i = 0; i < n; ++i) {
    const double *x = &in[(size_t)i * in_w];
    double *y = &out[(size_t)i * out_w];
    if (fn == "roe") {
      roe.compute_flux(x, x + 5, y, x + 10, x + 13, x + 16);
    } else if (fn == "viscous") {
      double g[5][3];
      memcpy(g, x, sizeof g);
      visc.compute_flux(g, x + 15, x + 20, y);
    } else if (fn == "primitives") {
      ComputePrimitives<Device>(x, y);
    } else if (fn == "venkat") {
      y[0] = VenkatLimiter<Device>::limit(x[0], x[1], x[2], x[3]);
    } else {
      y[0] = VanAlbadaLimiter<Device>::limit(x[0], x[1], x[2]);
    }
  }
  f = fopen(argv[3], "wb");
  if (!f) return perror(argv[3]), 1;
  fwrite(out.data(), 8, out.size(), f);
  fclose(f);
  return 0;
}
