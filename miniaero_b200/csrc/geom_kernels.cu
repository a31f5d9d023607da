// Device-side mesh geometry for the structured path: the reference's Face.C (area vector,
// tangent, binormal, centroid of a face) and MeshProcessor / ElementTopoHexa8 (cell centroid, 2x2x2 Gauss volume)
// evaluated on the GPU straight into the tile-packed structure-of-arrays layout, instead of being generated on the
// host as reference-format arrays, re-packed and uploaded (173 of the layout's ~260 bytes per cell).
//
// The arithmetic is mesh_geom.h — the same functions host_mesh.cpp uses — and this file is compiled with -fmad=false
// (IEEE division and square root are nvcc's default), so the device produces the host generator's bits, which are the
// reference's (tests/test_host_mesh.py, tests/test_gpu_structured.py).
#include <cuda_runtime.h>

#include <cstdint>

#include "geom_kernels.h"

namespace ma {

namespace {

// tile face j of tile T: face_code = (elem1 cell in the block's (n+2)^3 lattice) * 8 + elem1 local face
__global__ void __launch_bounds__(128) face_geometry_kernel(GridGen g, const TileInfoDev *__restrict__ tiles,
                                                            const uint32_t *__restrict__ face_code,
                                                            double *__restrict__ geom, long n_tile_faces, int with_tangents) {
  const TileInfoDev T = tiles[blockIdx.x];
  const long ly = g.b.n[1] + 2, lz = g.b.n[2] + 2;
  const size_t fcp = (size_t)((T.face_count + 15) / 16 * 16);
  for (int e = threadIdx.x; e < T.face_count; e += blockDim.x) {
    const size_t j = (size_t)T.face_start + e;
    const uint32_t code = face_code[j];
    const long lat = (long)(code >> 3);
    const int f = (int)(code & 7u);
    const int ci = (int)(lat / (ly * lz)) - 1, cj = (int)(lat / lz % ly) - 1, ck = (int)(lat % lz) - 1;
    double x[3], n[3], t[3], b[3];
    g.face_geometry(ci, cj, ck, f, x, n, t, b);
    if (with_tangents) {  // STRICT: global SoA [12][n_tile_faces]
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        geom[(size_t)(0 + d) * n_tile_faces + j] = n[d];
        geom[(size_t)(3 + d) * n_tile_faces + j] = t[d];
        geom[(size_t)(6 + d) * n_tile_faces + j] = b[d];
        geom[(size_t)(9 + d) * n_tile_faces + j] = x[d];
      }
    } else {  // FAST: tile-blocked [tile][6][faces rounded up to 16]
      const size_t base = (size_t)6 * T.face_start + e;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        geom[base + (size_t)(0 + d) * fcp] = n[d];
        geom[base + (size_t)(3 + d) * fcp] = x[d];
      }
    }
  }
}

// renumbered cell c is the block's cell new2old[c] (owned cells k-fastest, then the x-, y-, z-ghost groups)
__global__ void cell_geometry_kernel(GridGen g, const int *__restrict__ new2old, long n_cells, int stride,
                                     double *__restrict__ xyz, double *__restrict__ vol) {
  const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= stride) return;
  double ctr[3] = {0.0, 0.0, 0.0}, v = 1.0;  // the padding cells of the stride hold what the host layout gives them
  if (c < n_cells) {
    int i, j, k;
    g.cell_ijk(new2old[c], i, j, k);
    g.cell_geometry(i, j, k, ctr, &v);
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) xyz[(size_t)d * stride + c] = ctr[d];
  vol[c] = v;
}

}  // namespace

cudaError_t launch_device_geometry(const GridGen &g_dev, const TileInfoDev *tiles, int n_tiles, const uint32_t *face_code,
                                   double *geom, long n_tile_faces, int geom_components, const int *new2old,
                                   long n_cells, int stride, double *xyz, double *vol, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(geom, 0, (size_t)geom_components * n_tile_faces * sizeof(double), st);
  if (e != cudaSuccess) return e;
  if (n_tiles > 0)
    face_geometry_kernel<<<n_tiles, 128, 0, st>>>(g_dev, tiles, face_code, geom, n_tile_faces, geom_components == 12);
  const int threads = 128;
  cell_geometry_kernel<<<(unsigned)((stride + threads - 1) / threads), threads, 0, st>>>(g_dev, new2old, n_cells, stride,
                                                                                        xyz, vol);
  return cudaGetLastError();
}

}  // namespace ma
