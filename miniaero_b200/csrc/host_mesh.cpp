// Host-side in-code hex mesh generator: the behaviour of the reference's Parallel3DMesh +
// MeshProcessor + Face + ElementTopoHexa8 (one-time setup, host C++ as in the reference), written
// as a direct structured -> unstructured generator.  It produces the SAME arrays, bit for bit, that
// Parallel3DMesh::fillMeshData hands to the solver (Parallel3DMesh.h:173-449), but in O(cells) time
// and memory, in parallel, without per-face heap objects, node-set lookups or list sorts:
//
//   reference                                   here
//   ---------                                   ----
//   create_faces (MeshProcessor.C:39-128)       a face is created by the lower-numbered of its two
//     node-hash matching, creation order         cells, in (cell, local face 0..5) order -> count,
//                                                prefix-sum, fill
//   Face ctor (Face.C:37-98)                    face_geometry(): same expressions, same order
//   compute_cell_volumes / _centroid            cell_geometry(): 2x2x2 Gauss sum of detJ, node mean
//     (MeshProcessor.C:130-171, ElementTopoHexa8.C:52-148)
//   delete_ghosted_faces / extract / organize   a boundary face belongs to the domain side it lies on;
//     (MeshProcessor.C:173-229)                  faces of ghost cells with no neighbour are never made
//   setupCommunication + O(n^2) id search       neighbour ranks and (rank, global id)-ordered lists
//     (Parallel3DMesh.C:306-431, .h:279-319)     follow from the block structure
//
// All floating-point expressions are evaluated in the reference's order; this file is compiled with
// -ffp-contract=off so that no FMA contraction changes a bit.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <new>
#include <string>
#include <vector>

#include "host_common.h"
#include "mesh_geom.h"
#include "miniaero_b200.h"

namespace {

using ma::arrange;
using ma::Block;

const int kFaceDir[6][3] = {{0, -1, 0}, {1, 0, 0}, {0, 1, 0}, {-1, 0, 0}, {0, 0, -1}, {0, 0, 1}};
const int kOpposite[6] = {2, 3, 0, 1, 5, 4};

// the block's generator plus the node-coordinate tables it points at
struct Gen : ma::GridGen {
  ma::GridTables tables;
  void build_tables(double lx, double ly, double lz, double tan_ramp) { tables.build(*this, lx, ly, lz, tan_ramp); }
};

struct FaceArrays {  // every entry is written by the generator's parallel loops: no value-initialisation (BigVec)
  ma::BigVec<double> coords, normal, tangent, binormal;
  ma::BigVec<int> conn, slot;
  void resize(size_t n) {
    coords.resize(3 * n);
    normal.resize(3 * n);
    tangent.resize(3 * n);
    binormal.resize(3 * n);
    conn.resize(2 * n);
    slot.resize(2 * n);
  }
  ma_faces view() const {
    ma_faces f;
    f.nfaces = (int)(conn.size() / 2);
    f.coordinates = coords.data();
    f.face_normal = normal.data();
    f.face_tangent = tangent.data();
    f.face_binormal = binormal.data();
    f.face_cell_conn = conn.data();
    f.cell_flux_index = slot.data();
    return f;
  }
};

}  // namespace

struct ma_mesh_storage {
  ma_mesh view;
  Block block;
  ma::BigVec<double> cell_xyz, cell_vol;
  ma::BigVec<int> global_ids;
  FaceArrays internal;
  FaceArrays bc[6];
  std::vector<int> send_count, recv_count, send_ids, recv_ids;
};

extern "C" {

void ma_options_default(ma_options *o) {
  // Options.h:59-69
  std::memset(o, 0, sizeof(*o));
  o->problem_type = 0;
  o->nx = o->ny = o->nz = 10;
  o->ntimesteps = 1;
  o->dt = 5e-8;
  o->output_results = 0;
  o->output_frequency = 10;
  o->second_order_space = 0;
  o->viscous = 0;
  o->lx = o->ly = o->lz = 1.0;
}

int ma_options_read(const char *path, ma_options *o) {
  // Options.h:73-101: nine whitespace-separated values; the reference only warns when the file is
  // missing and carries on with garbage — here that is an error.
  if (!path || !o) return ma_set_error(MA_ERR_INVALID, "ma_options_read: null argument");
  std::ifstream f(path);
  if (!f) return ma_set_error(MA_ERR_IO, std::string(path) + " does not exist.");
  ma_options_default(o);
  f >> o->problem_type;
  f >> o->lx >> o->ly >> o->lz >> o->angle;
  f >> o->nx >> o->ny >> o->nz;
  f >> o->ntimesteps;
  f >> o->dt;
  f >> o->output_results;
  f >> o->output_frequency;
  f >> o->second_order_space;
  f >> o->viscous;
  if (f.fail()) return ma_set_error(MA_ERR_IO, std::string(path) + ": expected 14 numeric fields (Options.h:91-99)");
  return MA_OK;
}

int ma_mesh_generate(const ma_options *opt, int rank, int num_ranks, ma_mesh_storage **out) {
  if (!opt || !out) return ma_set_error(MA_ERR_INVALID, "ma_mesh_generate: null argument");
  *out = nullptr;
  if (opt->nx <= 0 || opt->ny <= 0 || opt->nz <= 0)
    return ma_set_error(MA_ERR_INVALID, "ma_mesh_generate: nx, ny, nz must be positive");
  if (num_ranks < 1 || rank < 0 || rank >= num_ranks)
    return ma_set_error(MA_ERR_INVALID, "ma_mesh_generate: bad rank / num_ranks");
  ma_mesh_storage *m = nullptr;
  try {
    m = new ma_mesh_storage();
    Gen g;
    if (!arrange(g.b, opt->nx, opt->ny, opt->nz, rank, num_ranks)) {
      delete m;
      return ma_set_error(MA_ERR_INVALID, "MPI number of ranks must be a power of 2.");  // Parallel3DMesh.C:262-265
    }
    const Block &b = g.b;
    for (int d = 0; d < 3; ++d)
      if (b.n[d] < 1) {
        delete m;
        return ma_set_error(MA_ERR_INVALID, "ma_mesh_generate: more blocks than cells in a direction");
      }
    const double PI = 3.14159265;  // Parallel3DMesh.h:473
    g.build_tables(opt->lx, opt->ly, opt->lz, std::tan(opt->angle * PI / 180.0));
    g.set_block_counts();
    const long ncells = g.nowned + g.nghost;
    if (ncells * 3 > 2000000000L) {  // int32 face ids, as the reference
      delete m;
      return ma_set_error(MA_ERR_INVALID, "ma_mesh_generate: more than 2^31 faces on one block");
    }
    m->block = b;

    // ---- cells: centroid for all, volume for all (the reference computes owned volumes and MPI-copies
    // the owner's value into ghosts, Parallel3DMesh.h:421-446: the same bits, since node coordinates are
    // functions of global indices only)
    m->cell_xyz.resize(3 * (size_t)ncells);
    m->cell_vol.resize((size_t)ncells);
    m->global_ids.resize((size_t)ncells);
#pragma omp parallel for schedule(static)
    for (long c = 0; c < ncells; ++c) {
      int i, j, k;
      g.cell_ijk(c, i, j, k);
      g.cell_geometry(i, j, k, &m->cell_xyz[3 * c], &m->cell_vol[c]);
      m->global_ids[c] = g.global_id(i, j, k);
    }

    // ---- internal faces in creation order: cell c creates local face f when the neighbour across f
    // exists and has a larger id (MeshProcessor.C:54-120).  Boundary faces of ghost cells are dropped
    // (delete_ghosted_faces, MeshProcessor.C:173-186); ghost-ghost faces are kept, as in the reference.
    ma::BigVec<long> first;
    first.resize((size_t)ncells + 1);
    first[0] = 0;
#pragma omp parallel for schedule(static)
    for (long c = 0; c < ncells; ++c) {
      int i, j, k;
      g.cell_ijk(c, i, j, k);
      int cnt = 0;
      for (int f = 0; f < 6; ++f) {
        long nb = g.cell_id(i + kFaceDir[f][0], j + kFaceDir[f][1], k + kFaceDir[f][2]);
        if (nb > c) ++cnt;
      }
      first[c + 1] = cnt;
    }
    for (long c = 0; c < ncells; ++c) first[c + 1] += first[c];
    const long nint = first[ncells];
    m->internal.resize((size_t)nint);
#pragma omp parallel for schedule(static)
    for (long c = 0; c < ncells; ++c) {
      int i, j, k;
      g.cell_ijk(c, i, j, k);
      long w = first[c];
      for (int f = 0; f < 6; ++f) {
        long nb = g.cell_id(i + kFaceDir[f][0], j + kFaceDir[f][1], k + kFaceDir[f][2]);
        if (nb > c) {
          FaceArrays &F = m->internal;
          g.face_geometry(i, j, k, f, &F.coords[3 * w], &F.normal[3 * w], &F.tangent[3 * w], &F.binormal[3 * w]);
          F.conn[2 * w] = (int)c;
          F.conn[2 * w + 1] = (int)nb;
          F.slot[2 * w] = f;
          F.slot[2 * w + 1] = kOpposite[f];
          ++w;
        }
      }
    }

    // ---- boundary sets.  Reference copy order is top,bottom,right,left,front,back but the solver sees
    // them as [bottom, top, front, back, right, left] (Parallel3DMesh.h:382-396); each set lists its
    // faces in creation order = owned-cell order.
    struct Side {
      int f;     // local face
      int axis;  // fixed axis
      int hi;    // 0: index 0, 1: index n-1
    };
    const Side sides[6] = {{0, 1, 0}, {2, 1, 1}, {4, 2, 0}, {5, 2, 1}, {1, 0, 1}, {3, 0, 0}};
    for (int s = 0; s < 6; ++s) {
      const Side &sd = sides[s];
      const bool on_boundary = sd.hi ? (b.blk[sd.axis] == b.np[sd.axis] - 1) : (b.blk[sd.axis] == 0);
      long cnt = 0;
      int a1 = (sd.axis == 0) ? 1 : 0, a2 = (sd.axis == 2) ? 1 : 2;  // the two free axes, in i<j<k order
      if (on_boundary) cnt = (long)b.n[a1] * b.n[a2];
      m->bc[s].resize((size_t)cnt);
      if (!cnt) continue;
      FaceArrays &F = m->bc[s];
      const int fixed = sd.hi ? b.n[sd.axis] - 1 : 0;
#pragma omp parallel for schedule(static)
      for (long w = 0; w < cnt; ++w) {
        int idx[3];
        idx[sd.axis] = fixed;
        idx[a1] = (int)(w / b.n[a2]);
        idx[a2] = (int)(w % b.n[a2]);
        g.face_geometry(idx[0], idx[1], idx[2], sd.f, &F.coords[3 * w], &F.normal[3 * w], &F.tangent[3 * w],
                        &F.binormal[3 * w]);
        F.conn[2 * w] = (int)g.cell_id(idx[0], idx[1], idx[2]);
        F.conn[2 * w + 1] = -1;
        F.slot[2 * w] = sd.f;
        F.slot[2 * w + 1] = -1;  // Face ctor leaves elem2_flux_index as passed (-1 default of FaceData)
      }
    }

    // ---- ghost exchange lists: per neighbour rank (ascending), ordered by global id
    m->send_count.assign(num_ranks, 0);
    m->recv_count.assign(num_ranks, 0);
    if (num_ranks > 1) {
      struct Nb {
        int rank, axis, hi;
      };
      std::vector<Nb> nbs;
      for (int d = 0; d < 3; ++d)
        for (int hi = 0; hi < 2; ++hi) {
          if (!(hi ? b.ghi[d] : b.glo[d])) continue;
          int nb_blk[3] = {b.blk[0], b.blk[1], b.blk[2]};
          nb_blk[d] += hi ? 1 : -1;
          nbs.push_back({nb_blk[0] + b.np[0] * (nb_blk[1] + b.np[1] * nb_blk[2]), d, hi});
        }
      std::sort(nbs.begin(), nbs.end(), [](const Nb &a, const Nb &c) { return a.rank < c.rank; });
      for (const Nb &nb : nbs) {
        const int d = nb.axis;
        const int a1 = (d == 0) ? 1 : 0, a2 = (d == 2) ? 1 : 2;
        const long cnt = (long)b.n[a1] * b.n[a2];
        m->send_count[nb.rank] = (int)cnt;
        m->recv_count[nb.rank] = (int)cnt;
        for (long w = 0; w < cnt; ++w) {
          int idx[3];
          idx[a1] = (int)(w / b.n[a2]);
          idx[a2] = (int)(w % b.n[a2]);
          idx[d] = nb.hi ? b.n[d] - 1 : 0;
          m->send_ids.push_back((int)g.cell_id(idx[0], idx[1], idx[2]));
          idx[d] = nb.hi ? b.n[d] : -1;
          m->recv_ids.push_back((int)g.cell_id(idx[0], idx[1], idx[2]));
        }
      }
    }

    // ---- view
    ma_mesh &v = m->view;
    std::memset(&v, 0, sizeof(v));
    v.num_owned_cells = (int)g.nowned;
    v.num_ghosts = (int)g.nghost;
    v.cell_coordinates = m->cell_xyz.data();
    v.cell_volumes = m->cell_vol.data();
    v.internal_faces = m->internal.view();
    v.num_boundary_sets = 6;
    const int pt = opt->problem_type;
    // Parallel3DMesh.h:382-396
    v.boundary_type[0] = (pt == 1) ? MA_BC_NOSLIP : MA_BC_TANGENT;       // bottom
    v.boundary_type[1] = (pt == 1) ? MA_BC_EXTRAPOLATE : MA_BC_TANGENT;  // top
    v.boundary_type[2] = MA_BC_TANGENT;                                  // front
    v.boundary_type[3] = MA_BC_TANGENT;                                  // back
    v.boundary_type[4] = MA_BC_EXTRAPOLATE;                              // right
    v.boundary_type[5] = (pt == 0) ? MA_BC_EXTRAPOLATE : MA_BC_INFLOW;   // left
    for (int s = 0; s < 6; ++s) v.boundary_faces[s] = m->bc[s].view();
    v.num_ranks = num_ranks;
    v.my_rank = rank;
    v.send_count = m->send_count.data();
    v.recv_count = m->recv_count.data();
    v.send_local_ids = m->send_ids.data();
    v.recv_local_ids = m->recv_ids.data();
  } catch (const std::bad_alloc &) {
    delete m;
    return ma_set_error(MA_ERR_NOMEM, "ma_mesh_generate: out of host memory");
  }
  *out = m;
  return MA_OK;
}

const ma_mesh *ma_mesh_view(const ma_mesh_storage *m) { return m ? &m->view : nullptr; }
const int *ma_mesh_global_ids(const ma_mesh_storage *m) { return m ? m->global_ids.data() : nullptr; }
void ma_mesh_decomposition(const ma_mesh_storage *m, int nproc[3], int block[3], int nlocal[3], int offset[3]) {
  if (!m) return;
  for (int d = 0; d < 3; ++d) {
    if (nproc) nproc[d] = m->block.np[d];
    if (block) block[d] = m->block.blk[d];
    if (nlocal) nlocal[d] = m->block.n[d];
    if (offset) offset[d] = m->block.off[d];
  }
}
void ma_mesh_free(ma_mesh_storage *m) { delete m; }

int ma_block_decomposition(const ma_options *opt, int rank, int num_ranks, int nproc[3], int block[3], int nlocal[3],
                           int offset[3]) {
  if (!opt || num_ranks < 1 || rank < 0 || rank >= num_ranks)
    return ma_set_error(MA_ERR_INVALID, "ma_block_decomposition: bad argument");
  Block b;
  if (!arrange(b, opt->nx, opt->ny, opt->nz, rank, num_ranks))
    return ma_set_error(MA_ERR_INVALID, "MPI number of ranks must be a power of 2.");  // Parallel3DMesh.C:262-265
  for (int d = 0; d < 3; ++d) {
    if (nproc) nproc[d] = b.np[d];
    if (block) block[d] = b.blk[d];
    if (nlocal) nlocal[d] = b.n[d];
    if (offset) offset[d] = b.off[d];
  }
  return MA_OK;
}

int ma_write_results(const char *path, const ma_mesh *mesh, const double *solution, int precision) {
  // TimeSolverExplicitRK4.h:514-538
  if (!path || !mesh || !solution) return ma_set_error(MA_ERR_INVALID, "ma_write_results: null argument");
  std::ofstream f(path, std::ios::out);
  if (!f) return ma_set_error(MA_ERR_IO, std::string("cannot open ") + path);
  if (precision > 0) f.precision(precision);
  for (int i = 0; i < mesh->num_owned_cells; ++i) {
    f << mesh->cell_coordinates[3 * (size_t)i] << "\t";
    f << mesh->cell_coordinates[3 * (size_t)i + 1] << "\t";
    f << mesh->cell_coordinates[3 * (size_t)i + 2] << "\t";
    for (int c = 0; c < 5; ++c) f << solution[5 * (size_t)i + c] << "\t";
    f << "\n";
  }
  return f.good() ? MA_OK : ma_set_error(MA_ERR_IO, std::string("write failed: ") + path);
}

}  // extern "C"
