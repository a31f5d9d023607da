// TEST INFRASTRUCTURE: see Kokkos_Core.hpp in this directory (Kokkos stand-in for the reference build).
#pragma once
#include <Kokkos_Core.hpp>
