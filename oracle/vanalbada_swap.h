// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Force-included (g++ -include) when oracle/build_ref.sh builds miniAero.cell.vanalbada: the UNMODIFIED reference
// sources with their stencil limiter switched from VenkatLimiter to the alternative the reference ships but never
// calls, VanAlbadaLimiter (VanAlbadaLimiter.h:45-65; Flux.h:36 includes it).  A reference maintainer would edit the two
// call sites StencilLimiter.h:455,459; here the class NAME used there is redirected by the preprocessor instead, so no
// reference line is touched: VenkatLimiter.h is included first (its include guard keeps the real class from being
// renamed), then every later `VenkatLimiter<Device>::limit(dumax, dumin, du, deltax3)` resolves to the adapter below.
#pragma once
#include <Kokkos_Core.hpp>

#include "VenkatLimiter.h"
#include "VanAlbadaLimiter.h"

template <class DeviceType>
class VanAlbadaAsStencilLimiter {
 public:
  KOKKOS_INLINE_FUNCTION static double limit(double dumax, double dumin, double du, double /*deltax3*/) {
    return VanAlbadaLimiter<DeviceType>::limit(dumax, dumin, du);
  }
};
#define VenkatLimiter VanAlbadaAsStencilLimiter
