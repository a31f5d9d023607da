// Launch interface of the CUDA kernels (kernels.cu is compiled twice: namespace ma_fast with FMA
// contraction, namespace ma_strict with -fmad=false; the solver picks one per ma_arith).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace ma {

struct TileInfoDev {
  int cell_start, cell_count, face_start, face_count;
  int cut_start;  // faces [0, cut_start) are closed or boundary faces, [cut_start, face_count) cut faces
  int n_eval;     // faces [0, n_eval) are evaluated by this tile's CTA; [n_eval, face_count) are imported (layout.h)
  int imp_area;   // the tile's import area in DevMesh::cut_flux, -1: none
  int pad_;
};

// Device mesh in the tile-packed structure-of-arrays layout (see DESIGN.md "Data layout in HBM").
struct DevMesh {
  int n_owned, n_cells, stride;        // cells; SoA component stride
  int n_tiles, slot_stride;
  long n_tile_faces;                   // component stride of face_geom
  int flux_smem_stride;                // per-component stride of the shared-memory face staging (>= faces of a tile)
  int rk_smem_stride;                  // per-component stride of the staged RK operands (>= cells of a tile)
  int grad_variant, flux_variant;      // FAST kernels: 0 = gather kernels, 1 = bulk-copy (TMA) staged tile kernels
  int limiter;                         // ma_limiter: 0 Venkatakrishnan, 1 Van Albada
  int tile_class;                      // capacity class of the staged tile kernels (kernels.cu: TileClass), -1: none fits
  int max_tile_cells, max_tile_faces, max_tile_halo, max_tile_local;
  int halo_stride;                     // tile_halo entries per tile (tile k's list starts at k * halo_stride, -1 padded)
  const TileInfoDev *tiles;
  const double *cell_xyz;              // [3][stride]
  const double *cell_vol;              // [stride]
  const uint16_t *slot_face;           // [6][slot_stride]
  const uint16_t *slot_nbr;            // [6][slot_stride] staged position of the cell across each slot (FAST)
  const double *face_geom;             // STRICT [12][n_tile_faces]: normal, tangent, binormal, centroid;
                                       // FAST tile-blocked [tile][6][faces of the tile rounded up to 16]: normal, centroid
  const int *face_left, *face_right;   // [n_tile_faces] renumbered cell ids (STRICT kernels only)
  const uint32_t *face_lr;             // [n_tile_faces] tile-local left | right << 16 (boundary: 0xFFFF - type)
  const int *tile_halo;                // outside cell of every cut face, halo_stride entries per tile
  // shared cut faces (layout.h): where each evaluated cut face publishes its flux (parallel to tile_halo, -1: nowhere),
  // the exchange buffer [areas][5][import_capacity], and its component stride; null / 0 when every tile evaluates all
  // of its faces
  const int *tile_pub;
  double *cut_flux;
  int import_capacity;
  double inflow[5];                    // TimeSolverExplicitRK4.h:218-223
};

// Shared cut faces order a tile group as [first flux pass | second pass] — the two colours of the tile lattice's
// checkerboard, each in Morton order — so the i-th tile of either pass is one half of the i-th Morton pair.  The
// gradient sweep has no passes: it walks the group pair by pair (first[0], second[0], first[1], second[1], ...), i.e.
// in (nearly) the Morton order of the whole group, and the neighbour cells it gathers are again the ones the tiles
// just before it staged.  n_first == ntiles (or 0): plain order.  A permutation of [0, ntiles) for every n_first
// (tools/layout_check.cpp checks it).
#ifdef __CUDACC__
__host__ __device__
#endif
inline int interleaved_tile(int b, int n_first, int ntiles) {
  const int n_second = ntiles - n_first;
  const int paired = n_first < n_second ? n_first : n_second;
  if (b < 2 * paired) return (b & 1) ? n_first + (b >> 1) : (b >> 1);
  const int rem = b - 2 * paired;
  return n_first > n_second ? paired + rem : n_first + paired + rem;
}

struct StageArgs {
  const double *V;     // primitives (rho,u,v,w,T) of the stage state, every cell a tile touches ("solution_temp")
  double *Vnext;       // primitives of the next stage state (for the last stage: of the new solution)
  double *Un;          // conservative state at the start of the step ("solution_n"); written by the last stage
  double *Acc;         // running RK sum ("solution_np1"), updated in place
  const double *grad;  // [15][stride]
  const double *lim;   // [5][stride]
  double dt, alpha_next, beta;
  int kind;            // 0 first stage (stage state == Un), 1 middle, 2 last
};

}  // namespace ma

#define MA_DECLARE_KERNEL_API(NS)                                                                                    \
  namespace NS {                                                                                                     \
  /* GreenGauss.h:51-270 + StencilLimiter.h:56-500 fused, cell-centric, tiles [tile_begin, tile_begin+ntiles) */     \
  /* n_first: tiles of the range that belong to the first flux pass (shared cut faces; == ntiles otherwise) */       \
  cudaError_t launch_grad_limiter(const ma::DevMesh &m, const double *V, double *grad, double *lim, bool second,     \
                                  int tile_begin, int ntiles, int n_first, int threads, cudaStream_t st);            \
  /* Flux.h:52-229 + the four *_BC.h + TimeSolverExplicitRK4.h:106-128 fused */                                      \
  cudaError_t launch_flux_rk(const ma::DevMesh &m, const ma::StageArgs &a, bool second, bool viscous,                \
                             int tile_begin, int ntiles, int threads, cudaStream_t st);                              \
  cudaError_t flux_rk_prepare(const ma::DevMesh &m, int smem_bytes);                                                                     \
  /* shared memory per CTA of the two stage kernels for this mesh */                                                 \
  size_t grad_smem_bytes(const ma::DevMesh &m, bool second);                                                         \
  size_t flux_smem_bytes(const ma::DevMesh &m, bool second, bool viscous);                                           \
  /* smallest capacity class of the staged tile kernels that holds every tile of the mesh, or -1 */                  \
  int pick_tile_class(int max_cells, int max_faces, int max_halo);                                                   \
  /* threads per CTA the staged kernels of that class want (0: caller's choice) */                                   \
  int tile_class_threads(int tile_class, int which);                                                                 \
  /* GasModel.h:70-90 over the owned cells: conservative Un -> primitives V */                                       \
  cudaError_t launch_primitives(const ma::DevMesh &m, const double *Un, double *V, cudaStream_t st);                 \
  /* caller-order AoS conservative state -> Un (renumbered SoA) and its primitives V in one pass */                  \
  cudaError_t launch_set_state(const ma::DevMesh &m, const double *aos, const int *old2new, double *Un, double *V,   \
                               cudaStream_t st);                                                                     \
  /* Initial_Conditions.h:38-133 */                                                                                  \
  cudaError_t launch_initial_conditions(const ma::DevMesh &m, double *Un, int problem_type, double midx,             \
                                        cudaStream_t st);                                                            \
  cudaError_t probe_roe(int n, const double *vl, const double *vr, const double *nn, const double *tt,               \
                        const double *bb, double *flux, cudaStream_t st);                                            \
  cudaError_t probe_viscous(int n, const double *g, const double *v, const double *a, double *vf, cudaStream_t st);  \
  cudaError_t probe_primitives(int n, const double *u, double *v, cudaStream_t st);                                  \
  cudaError_t probe_venkat(int n, const double *dmax, const double *dmin, const double *du, const double *dx3,       \
                           double *phi, cudaStream_t st);                                                            \
  cudaError_t probe_vanalbada(int n, const double *dmax, const double *dmin, const double *du, double *phi,          \
                              cudaStream_t st);                                                                      \
  }

MA_DECLARE_KERNEL_API(ma_fast)
MA_DECLARE_KERNEL_API(ma_strict)
