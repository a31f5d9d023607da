// Device-side layout topology for the structured path (second half): cell renumbering, slot
// maps, tile-local connectivity, face codes, outside-cell lists and publish slots are stamped on the GPU, one CTA per
// tile, from the per-pattern templates of a TopoPlan — what Parallel3DMesh::fillMeshData + MeshProcessor::create_faces
// (Parallel3DMesh.h:173-449, MeshProcessor.C:39-128) and the host layout builder produce with O(cells) host loops.
// The arithmetic is topology_stamp.h, shared with the host-side check tools/topology_compare.cpp.
#include "topology_kernels.h"

namespace ma {

namespace {

__global__ void __launch_bounds__(128) topo_cells_kernel(const TopoView t) {
  const int k = blockIdx.x;
  const int n = t.tiles[k].cell_count;
  for (int lc = threadIdx.x; lc < n; lc += blockDim.x) topo_stamp_cell(t, k, lc);
}
__global__ void topo_ghost_ids_kernel(int *__restrict__ new2old, int *__restrict__ old2new, int n_owned, long n_cells) {
  const long c = n_owned + (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_cells) new2old[c] = old2new[c] = (int)c;  // ghosts keep their ids (appended after the owned cells)
}
__global__ void __launch_bounds__(128) topo_faces_kernel(const TopoView t) {
  const int k = blockIdx.x;
  const int n = t.tiles[k].face_count;
  for (int e = threadIdx.x; e < n; e += blockDim.x) topo_stamp_face(t, k, e);
}
__global__ void __launch_bounds__(128) topo_pub_kernel(const TopoView t) {
  const int k = blockIdx.x;
  const int n = t.tiles[k].n_eval - t.tiles[k].cut_start;
  for (int q = threadIdx.x; q < n; q += blockDim.x) topo_stamp_pub(t, k, q);
}

}  // namespace

cudaError_t launch_device_topology(const TopoView &t, long n_cells, cudaStream_t st) {
  if (t.n_tiles <= 0) return cudaSuccess;
  topo_cells_kernel<<<t.n_tiles, 128, 0, st>>>(t);
  const long ng = n_cells - t.n_owned;
  if (ng > 0) topo_ghost_ids_kernel<<<(unsigned)((ng + 255) / 256), 256, 0, st>>>(t.new2old, t.old2new, t.n_owned, n_cells);
  topo_faces_kernel<<<t.n_tiles, 128, 0, st>>>(t);
  if (t.tile_pub) topo_pub_kernel<<<t.n_tiles, 128, 0, st>>>(t);  // reads the slot maps of OTHER tiles: after the face kernel
  return cudaGetLastError();
}

}  // namespace ma
