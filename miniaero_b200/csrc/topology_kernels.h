// Device-side construction of the O(cells) layout arrays of a structured block from a TopoPlan (layout.h): see
// topology_stamp.h for what is computed and topology_kernels.cu for the launches.
#pragma once
#include <cuda_runtime.h>

#include "topology_stamp.h"

namespace ma {

// every pointer of `t` is a device pointer; the output arrays must be initialised (slot_face 0, slot_nbr 0xFFFF,
// face_lr / face_code 0, tile_halo / tile_pub -1).  n_cells = owned + ghost cells.
cudaError_t launch_device_topology(const TopoView &t, long n_cells, cudaStream_t st);

}  // namespace ma
