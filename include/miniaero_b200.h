/*
 * miniaero_b200.h — C ABI of the B200-native miniAero explicit-RK4 finite-volume step.
 *
 * This is the drop-in boundary for the reference's only seam on the hot path,
 *
 *     TimeSolverExplicitRK4<Device>(MeshData<Device>&, const Options&);   TimeSolverExplicitRK4.h:164,180-196
 *     void Solve();                                                       TimeSolverExplicitRK4.h:166,207
 *     call site: Main.C:139-141
 *
 * expressed as `extern "C"` functions over plain pointers and sizes.  The structs below are
 * field-for-field restatements of the reference's host-visible containers (file:line given on
 * each), with Kokkos::View handles replaced by raw row-major (LayoutRight) host pointers, which
 * is the layout the reference's host mirrors have (ViewTypes.h:41, Faces.h:88-123).
 *
 * Conventions
 *   - every function returns MA_OK (0) or a negative ma_status; `ma_last_error()` gives the text
 *     of the most recent failure on the calling thread.  Nothing throws across the boundary.
 *   - input arrays are borrowed only for the duration of the call that takes them; the library
 *     owns all device memory.  Output arrays are caller-allocated host buffers.
 *   - there is NO CPU fallback: every solver entry point fails with MA_ERR_CUDA when no sm_100
 *     class device is usable.
 *   - one solver object is driven by one host thread at a time (as the reference: Main.C:139-141).
 */
#ifndef MINIAERO_B200_H_
#define MINIAERO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MA_ABI_VERSION 3
#define MA_MAX_BC_SETS 16

typedef enum ma_status {
  MA_OK = 0,
  MA_ERR_INVALID = -1, /* bad argument / inconsistent mesh */
  MA_ERR_CUDA = -2,    /* CUDA runtime or kernel failure, or no usable device */
  MA_ERR_NCCL = -3,    /* NCCL failure or libnccl not loadable */
  MA_ERR_IO = -4,      /* file could not be read / written */
  MA_ERR_NOMEM = -5
} ma_status;

/* Options.h:47-57 — the nine whitespace-separated values of ./miniaero.inp, in file order
 * (Options.h:91-99): problem_type / lx ly lz angle / nx ny nz / ntimesteps / dt / output_results /
 * output_frequency / second_order_space / viscous. */
typedef struct ma_options {
  int problem_type; /* 0 Sod, 1 viscous flat plate, 2 inviscid ramp */
  double lx, ly, lz, angle;
  int nx, ny, nz;
  int ntimesteps;
  double dt;
  int output_results;
  int output_frequency;
  int second_order_space;
  int viscous;
} ma_options;

/* Faces.h:42-71 (struct Faces<Device>), host-mirror layout of copy_faces (Faces.h:88-123). */
typedef struct ma_faces {
  int nfaces;
  const double *coordinates;   /* [nfaces][3] face centroid                       Faces.h:61 */
  const double *face_normal;   /* [nfaces][3] area-weighted normal, elem1 -> elem2 Faces.h:62 */
  const double *face_tangent;  /* [nfaces][3] unit tangent                         Faces.h:63 */
  const double *face_binormal; /* [nfaces][3] normal x tangent (area magnitude)    Faces.h:64 */
  const int *face_cell_conn;   /* [nfaces][2] elem1, elem2 (elem2 ignored for boundary sets) Faces.h:65 */
  const int *cell_flux_index;  /* [nfaces][2] slot (0..5) of this face in elem1 / elem2      Faces.h:66 */
} ma_faces;

/* Boundary-set names of MeshData::boundary_faces (Parallel3DMesh.h:382-396,
 * TimeSolverExplicitRK4.h:236-249): "Extrapolate" | "Tangent" | "Inflow" | "NoSlip". */
typedef enum ma_bc_type { MA_BC_EXTRAPOLATE = 0, MA_BC_TANGENT = 1, MA_BC_INFLOW = 2, MA_BC_NOSLIP = 3 } ma_bc_type;

/* MeshData.h:43-58 (struct MeshData<Device>) + Cells.h:41-76 (coordinates_, volumes_). */
typedef struct ma_mesh {
  int num_owned_cells;            /* MeshData.h:46 */
  int num_ghosts;                 /* MeshData.h:45; ghosts are cells [num_owned_cells, num_owned_cells+num_ghosts) */
  const double *cell_coordinates; /* [ncells][3]  Cells.h:63 */
  const double *cell_volumes;     /* [ncells]     Cells.h:64 */
  ma_faces internal_faces;        /* MeshData.h:56 */
  int num_boundary_sets;          /* MeshData.h:57 — order is the order gradients/limiters walk them */
  int boundary_type[MA_MAX_BC_SETS]; /* ma_bc_type of each set */
  ma_faces boundary_faces[MA_MAX_BC_SETS];
  /* ghost exchange lists (MeshData.h:50-55); all NULL / 0 for a single-domain run */
  int num_ranks;
  int my_rank;
  const int *send_count;     /* [num_ranks] cells sent to each rank      MeshData.h:51 */
  const int *recv_count;     /* [num_ranks] ghosts received from each rank MeshData.h:51 */
  const int *send_local_ids; /* [sum send_count] ordered by (rank, global id) Parallel3DMesh.h:290-319 */
  const int *recv_local_ids; /* [sum recv_count]                               */
} ma_mesh;

/* ---- error text ------------------------------------------------------------------------- */
const char *ma_last_error(void);
int ma_abi_version(void);

/* ---- Options (host) — replaces Options::read_options_file, Options.h:73-101 ---------------- */
void ma_options_default(ma_options *opt); /* Options.h:59-69 (viscous additionally defaults to 0) */
int ma_options_read(const char *path, ma_options *opt);

/* ---- In-code hex mesh (host) — replaces Parallel3DMesh + MeshProcessor + Face + ElementTopoHexa8
 * (Parallel3DMesh.h:173-449, MeshProcessor.C:39-229, Face.C:37-98, ElementTopoHexa8.C:134-148).
 * Produces, for block `rank` of `num_ranks` (2^k, Parallel3DMesh.C:247-303), the same arrays the
 * reference hands to the solver; internal faces are in creation order (the reference then applies
 * an unseeded std::random_shuffle, Parallel3DMesh.h:362, which no result depends on). */
typedef struct ma_mesh_storage ma_mesh_storage;
int ma_mesh_generate(const ma_options *opt, int rank, int num_ranks, ma_mesh_storage **out);
const ma_mesh *ma_mesh_view(const ma_mesh_storage *m);
/* global element id of each local cell, [ncells] (Parallel3DMesh.h:462-464) */
const int *ma_mesh_global_ids(const ma_mesh_storage *m);
/* block decomposition of this rank: nproc[3], block[3], local n[3], offset[3] (Parallel3DMesh.C:247-303) */
void ma_mesh_decomposition(const ma_mesh_storage *m, int nproc[3], int block[3], int nlocal[3], int offset[3]);
void ma_mesh_free(ma_mesh_storage *m);
/* The block decomposition alone (Parallel3DMesh.C:247-303: 2^k ranks, bisect the largest remaining dimension):
 * blocks per direction, this rank's block coordinates, its local cell counts and global offsets. */
int ma_block_decomposition(const ma_options *opt, int rank, int num_ranks, int nproc[3], int block[3], int nlocal[3],
                           int offset[3]);

/* ---- Halo communicator (NCCL send/recv over NVLink) — replaces MPI_COMM_WORLD as used by
 * communicate_ghosted_cell_data, CopyGhost.C:41-79.  One process per GPU; rank 0 obtains the id
 * and distributes its MA_COMM_ID_BYTES bytes by any out-of-band channel (torch.distributed
 * broadcast, a file, ...). */
#define MA_COMM_ID_BYTES 128
typedef struct ma_comm ma_comm;
int ma_comm_get_unique_id(unsigned char id[MA_COMM_ID_BYTES]);
int ma_comm_create(const unsigned char id[MA_COMM_ID_BYTES], int num_ranks, int rank, int device, ma_comm **out);
void ma_comm_destroy(ma_comm *c);

/* ---- Solver — replaces TimeSolverExplicitRK4<Device> ------------------------------------------ */
typedef enum ma_arith {
  MA_ARITH_FAST = 0,  /* FMA contraction, Newton reciprocals / square roots, algebraically regrouped Roe dissipation
                         and limiter; the production path.  Assumes each face's (normal, tangent, binormal) is an
                         orthogonal frame with unit tangent (as Face.C:81-96 builds it): checked at create time */
  MA_ARITH_STRICT = 1 /* IEEE-754 evaluation in the reference's source order (no FMA): bit-for-bit
                         comparable with the reference's -DCELL_FLUX build */
} ma_arith;

typedef enum ma_limiter {
  MA_LIMITER_VENKAT = 0,    /* VenkatLimiter.h:45-73 — what StencilLimiter.h:455,459 call */
  MA_LIMITER_VANALBADA = 1  /* VanAlbadaLimiter.h:45-65 — shipped by the reference (Flux.h:36) but never called; a
                               maintainer switches by editing those two call sites.  Runs in the same staged
                               gradient kernel as the default limiter (IEEE divisions per face: not tuned) */
} ma_limiter;

typedef struct ma_solver_config {
  int device;       /* CUDA device ordinal */
  int arith;        /* ma_arith */
  int tile_dims[3]; /* cells per tile along the three mesh directions; 0 = library default */
  int block_threads; /* threads per CTA of the flux kernel; 0 = library default */
  ma_comm *comm;    /* NULL for a single-domain run; required when mesh->num_ghosts > 0 */
  int overlap_halo; /* non-zero: run interior tiles while the halo exchange is in flight */
  void *stream;     /* cudaStream_t to run on; NULL = a stream owned by the solver */
  int limiter;      /* ma_limiter (second-order runs) */
  int share_cut_faces; /* FAST staged flux kernel: 0 = library default (on for meshes of at least a few thousand tiles),
                          1 = every face between two tiles is evaluated by ONE of them (the tile of the earlier of two
                          flux launches per stage — the checkerboard colours of the tile lattice) and its flux handed to
                          the other through an exchange buffer, -1 = both tiles evaluate it (one flux launch per stage).
                          Same results bit for bit either way. */
} ma_solver_config;
void ma_solver_config_default(ma_solver_config *cfg);

typedef struct ma_solver ma_solver;

/* Constructor (TimeSolverExplicitRK4.h:180-196).  Copies the mesh, renumbers cells into tiles,
 * converts to the device structure-of-arrays layout and uploads it.  The permutation is kept so
 * that every get/set call below speaks the caller's original cell order. */
int ma_solver_create(const ma_mesh *mesh, const ma_options *opt, const ma_solver_config *cfg, ma_solver **out);
/* Parallel3DMesh + MeshProcessor + the constructor in one call, for the in-code mesh the reference always runs on
 * (Main.C:113-141: Parallel3DMesh(nx, ny, nz, lx, ly, lz, problem_type, angle) -> fillMeshData -> TimeSolverExplicitRK4):
 * rank `rank` of `num_ranks` builds its block's device layout straight from (i, j, k) — the reference-format face and
 * cell arrays are never materialised — and evaluates Face.C's face geometry and ElementTopoHexa8's cell volumes ON THE
 * DEVICE.  The result is the solver ma_solver_create(ma_mesh_generate(opt, rank, num_ranks), ...) gives, bit for bit,
 * for a third of the set-up time and host memory; get/set calls speak the block's reference cell order. */
int ma_solver_create_structured(const ma_options *opt, int rank, int num_ranks, const ma_solver_config *cfg,
                                ma_solver **out);
/* Owned and ghost cells of the solver's block (either may be NULL). */
int ma_solver_num_cells(const ma_solver *s, int *owned, int *ghosts);
void ma_solver_destroy(ma_solver *s);

/* Initial conditions of Solve() (TimeSolverExplicitRK4.h:324-338): Sod states split at lx/2 for
 * problem_type 0, the inflow state everywhere otherwise. */
int ma_solver_initialize(ma_solver *s);
/* Advance `nsteps` RK4 time steps (the loop body TimeSolverExplicitRK4.h:340-491). */
int ma_solver_step(ma_solver *s, int nsteps);
/* Solve() == initialize + step(opt.ntimesteps) + optional progress lines (…RK4.h:207-496). */
int ma_solver_solve(ma_solver *s);
/* Block the host until all queued work of the solver is finished. */
int ma_solver_synchronize(ma_solver *s);

/* Conserved variables [num_owned_cells][5] (rho, rho u, rho v, rho w, rho E), caller's cell order —
 * what Solve() copies back as "solution_n" (TimeSolverExplicitRK4.h:516).  `host` may be pinned. */
int ma_solver_get_solution(ma_solver *s, double *host);
int ma_solver_set_solution(ma_solver *s, const double *host);
/* Asynchronous ensemble member: upload `state_in` ([num_owned_cells][5], caller's cell order), advance `nsteps` RK4
 * steps from it, download the new state into `state_out`.  Returns once the work is queued; `state_in` must stay
 * untouched and `state_out` is valid only after ma_solver_synchronize().  Consecutive submissions are pipelined
 * over three streams (upload of member i+1 and download of member i-1 overlap the stepping of member i), which
 * needs page-locked host buffers to be effective.  The members are independent: each starts from its own
 * `state_in`, as successive Solve() calls of the reference on different initial states would
 * (TimeSolverExplicitRK4.h:207-539 with the state of :324-338 replaced by the caller's). */
int ma_solver_submit(ma_solver *s, const double *state_in, double *state_out, int nsteps);

/* Intermediate fields of the most recent RK stage, caller's cell order (parity checks):
 *   MA_FIELD_GRADIENT [num_owned_cells][5][3]  GreenGauss.h:228-270
 *   MA_FIELD_LIMITER  [num_owned_cells][5]     StencilLimiter.h:319-350
 *   MA_FIELD_STAGE_PRIMITIVES [num_owned_cells][5]  (rho,u,v,w,T) = ComputePrimitives (GasModel.h:70-90) of
 *                        "solution_temp" (TimeSolverExplicitRK4.h:355) — the form the stage state is kept in */
typedef enum ma_field { MA_FIELD_GRADIENT = 0, MA_FIELD_LIMITER = 1, MA_FIELD_STAGE_PRIMITIVES = 2 } ma_field;
int ma_solver_get_field(ma_solver *s, int field, double *host);

typedef struct ma_timing {
  double step_seconds;      /* device time of all ma_solver_step calls so far (CUDA events) */
  long long steps;          /* RK4 time steps taken */
  long long cell_updates;   /* owned cells x steps */
  double grad_seconds;      /* share spent in the gradient+limiter kernel (events; 0 when not profiled) */
  double flux_seconds;      /* share spent in the flux+gather+RK kernel */
  double halo_seconds;      /* share spent in pack/exchange/unpack */
  long long kernel_launches; /* kernels launched by ma_solver_step so far */
  size_t device_bytes;      /* device memory held by the solver */
  int num_tiles;
  int tile_faces_total;     /* faces summed over tiles (each tile-boundary face counted twice) */
  /* block-decomposed runs (zero otherwise) */
  int num_interior_tiles;   /* tiles that touch no ghost cell: they run while the halo exchange is in flight */
  int num_send_cells, num_recv_cells; /* cells packed / ghosts unpacked per exchange */
  double halo_wait_seconds; /* time the compute stream sat waiting for an exchange before a boundary-tile launch
                               (events; 0 when not profiled): the part of halo_seconds that was NOT hidden */
  long long faces_evaluated; /* face fluxes evaluated per stage, summed over tiles (== tile_faces_total unless cut
                                faces are shared: then every face between two tiles counts once) */
} ma_timing;
int ma_solver_get_timing(ma_solver *s, ma_timing *t);
int ma_solver_reset_timing(ma_solver *s);
/* non-zero: bracket every kernel with events to fill grad/flux/halo_seconds (serialises streams) */
int ma_solver_set_profiling(ma_solver *s, int enabled);

/* Test hook: one of the solver's device-resident layout arrays, copied to `out` (at most `capacity` bytes; the array's
 * size in bytes is returned through `size` either way).  Names: "tiles", "slot_face", "slot_nbr", "face_lr",
 * "tile_halo", "tile_pub", "old2new", "face_geom", "cell_xyz", "cell_vol".  tests/test_gpu_topology.py uses it to
 * check the device-side topology builder against the host builder byte for byte. */
int ma_solver_debug_array(ma_solver *s, const char *name, void *out, size_t capacity, size_t *size);

/* results.<rank> writer of Solve() (TimeSolverExplicitRK4.h:514-538): x y z rho rhou rhov rhow rhoE,
 * tab separated, `precision` significant digits (the reference uses the ostream default, 6). */
int ma_write_results(const char *path, const ma_mesh *mesh, const double *solution, int precision);

/* Mantevo YAML report — the reference compiles YAML_Doc / YAML_Element (YAML_Doc.C:27-67, YAML_Element.C:97-104)
 * but never calls them; the host driver here does.  Same grammar: "Mini-Application Name/Version" header, then
 * "key: value" lines with two spaces of indentation per level, written to
 * <dir>/<name>-<version>_<YYYY:MM:DD-HH:MM:SS>.yaml (dir NULL or "" = "."). */
typedef struct ma_report {
  const char *app_name;    /* NULL = "miniAero-b200" */
  const char *app_version; /* NULL = "1.0" */
  const ma_options *options;
  int num_ranks;
  int blocks[3];            /* block decomposition (Parallel3DMesh.C:247-303) */
  long long global_cells;   /* owned cells summed over ranks */
  const ma_timing *timing;  /* of the reporting rank */
  double setup_seconds, run_seconds, total_seconds; /* Main.C "Setup time" / "Device Run time" / "Total elapsed time" */
  double hbm_peak_gbs;      /* measured HBM bandwidth the roofline fraction refers to; 0 = omit */
  const char *device_name;  /* NULL = queried from the CUDA runtime */
} ma_report;
/* Writes the file and, when path_out is not NULL, its path (truncated to path_len). */
int ma_write_yaml_report(const ma_report *report, const char *dir, char *path_out, size_t path_len);

/* ---- Device-function probes (unit parity tests of the physics, run on the GPU) ------------------
 * Each evaluates one reference device function for n independent inputs (arrays are host pointers,
 * row-major).  arith is ma_arith. */
/* Roe_Flux.h:49-265: primitives [n][5] x2, normal/tangent/binormal [n][3] -> flux [n][5] */
int ma_probe_roe_flux(int n, const double *prim_l, const double *prim_r, const double *normal,
                      const double *tangent, const double *binormal, double *flux, int arith, int device);
/* Viscous_Flux.h:65-98: grad [n][5][3], primitives [n][5], a_vec [n][3] -> vflux [n][5] */
int ma_probe_viscous_flux(int n, const double *grad, const double *prim, const double *normal, double *vflux,
                          int arith, int device);
/* GasModel.h:70-90: conservatives [n][5] -> primitives [n][5] */
int ma_probe_primitives(int n, const double *cons, double *prim, int arith, int device);
/* VenkatLimiter.h:45-73 / VanAlbadaLimiter.h:45-65: dumax, dumin, du, deltax3 [n] -> phi [n] */
int ma_probe_venkat(int n, const double *dumax, const double *dumin, const double *du, const double *deltax3,
                    double *phi, int arith, int device);
int ma_probe_vanalbada(int n, const double *dumax, const double *dumin, const double *du, double *phi, int arith,
                       int device);

#ifdef __cplusplus
}
#endif
#endif /* MINIAERO_B200_H_ */
