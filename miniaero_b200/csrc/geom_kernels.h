// Device-side geometry of the structured mesh (geom_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"
#include "mesh_geom.h"

namespace ma {

// g_dev: the block's generator with xs / ys / zs pointing at DEVICE copies of the node tables.  Fills
// geom ([6] tile-blocked or [12][n_tile_faces]), xyz[3][stride] and vol[stride] on stream st.
cudaError_t launch_device_geometry(const GridGen &g_dev, const TileInfoDev *tiles, int n_tiles, const uint32_t *face_code,
                                   double *geom, long n_tile_faces, int geom_components, const int *new2old,
                                   long n_cells, int stride, double *xyz, double *vol, cudaStream_t st);

}  // namespace ma
