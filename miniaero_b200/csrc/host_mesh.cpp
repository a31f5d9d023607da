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
#include "miniaero_b200.h"

namespace {

// Hex8 local face -> local nodes (MeshProcessor.C:44); slot s of a cell is local face s.
// 0: -y, 1: +x, 2: +y, 3: -x, 4: -z, 5: +z
const int kFaceNodes[6][4] = {{0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {0, 3, 2, 1}, {4, 5, 6, 7}};
const int kFaceDir[6][3] = {{0, -1, 0}, {1, 0, 0}, {0, 1, 0}, {-1, 0, 0}, {0, 0, -1}, {0, 0, 1}};
const int kOpposite[6] = {2, 3, 0, 1, 5, 4};
// node n of cell (i,j,k) is (i+di, j+dj, k+dk)  (Parallel3DMesh.C:83-95)
const int kNodeOff[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};

// 2x2x2 Gauss points (ElementTopoHexa8.h:35-43)
const double kG = 0.577350269189626;
const double kXi[8] = {-kG, kG, kG, -kG, -kG, kG, kG, -kG};
const double kEta[8] = {-kG, -kG, kG, kG, -kG, -kG, kG, kG};
const double kZeta[8] = {-kG, -kG, -kG, -kG, kG, kG, kG, kG};

struct Block {
  // Parallel3DMesh.C:247-303
  int nprocs = 1, rank = 0;
  int np[3] = {1, 1, 1};
  int blk[3] = {0, 0, 0};
  int gn[3] = {0, 0, 0};   // global cells
  int n[3] = {0, 0, 0};    // local cells
  int off[3] = {0, 0, 0};  // global offset of local (0,0,0)
  int glo[3] = {0, 0, 0};  // 1 if a ghost layer exists at index -1
  int ghi[3] = {0, 0, 0};  // 1 if a ghost layer exists at index n
};

bool arrange(Block &b, int gnx, int gny, int gnz, int rank, int nprocs) {
  b.nprocs = nprocs;
  b.rank = rank;
  b.gn[0] = gnx, b.gn[1] = gny, b.gn[2] = gnz;
  int left = nprocs;
  int t[3] = {gnx, gny, gnz};
  b.np[0] = b.np[1] = b.np[2] = 1;
  while (left != 1) {
    if (left % 2 != 0) return false;  // "MPI number of ranks must be a power of 2."
    left /= 2;
    int mx = t[0];
    if (t[1] > mx) mx = t[1];
    if (t[2] > mx) mx = t[2];
    for (int d = 0; d < 3; ++d) {
      if (t[d] == mx) {
        b.np[d] *= 2;
        t[d] = b.gn[d] / b.np[d];
        break;
      }
    }
  }
  for (int d = 0; d < 3; ++d) b.n[d] = t[d];
  b.blk[0] = rank % b.np[0];
  int rest = rank / b.np[0];
  b.blk[1] = rest % b.np[1];
  rest /= b.np[1];
  b.blk[2] = rest % b.np[2];
  for (int d = 0; d < 3; ++d) {
    b.off[d] = b.gn[d] / b.np[d] * b.blk[d];
    b.glo[d] = (b.np[d] != 1 && b.blk[d] != 0) ? 1 : 0;
    b.ghi[d] = (b.np[d] != 1 && b.blk[d] != b.np[d] - 1) ? 1 : 0;
  }
  return true;
}

struct Gen {
  Block b;
  double lx, ly, lz, tan_ramp;
  std::vector<double> xs, zs, ys;  // node coordinate tables over global index -1 .. gn+1
  int ysj = 0;
  long nowned = 0, nghost = 0;
  long base[3] = {0, 0, 0};  // first local id of the x-, y-, z-ghost groups
  int ng[3] = {0, 0, 0};     // ghost layers per direction (0..2)

  // Parallel3DMesh.h:466-487 — note the two different expressions for y either side of lx/2
  void build_tables() {
    const int NX = b.gn[0], NY = b.gn[1], NZ = b.gn[2];
    xs.resize(NX + 3);
    zs.resize(NZ + 3);
    for (int gi = -1; gi <= NX + 1; ++gi) xs[gi + 1] = (double)gi / (NX)*lx;
    for (int gk = -1; gk <= NZ + 1; ++gk) zs[gk + 1] = (double)gk / (NZ)*lz;
    ysj = NY + 3;
    ys.resize((size_t)(NX + 3) * ysj);
    for (int gi = -1; gi <= NX + 1; ++gi) {
      const double x = xs[gi + 1];
      for (int gj = -1; gj <= NY + 1; ++gj) {
        double y;
        if (x < lx / 2.0) {
          y = (double)gj / (NY)*ly;
        } else {
          double y_ramp = (x - (lx / 2.0)) * tan_ramp;
          double ly_scaled = ly - y_ramp;
          y = y_ramp + (double)gj * ly_scaled / (NY);
        }
        ys[(size_t)(gi + 1) * ysj + (gj + 1)] = y;
      }
    }
  }
  inline void node(int i, int j, int k, double *c) const {  // local node index -> coordinate
    const int gi = b.off[0] + i, gj = b.off[1] + j, gk = b.off[2] + k;
    c[0] = xs[gi + 1];
    c[1] = ys[(size_t)(gi + 1) * ysj + (gj + 1)];
    c[2] = zs[gk + 1];
  }

  // local cell id in the reference's numbering (Parallel3DMesh.C:80-174): owned cells k-fastest,
  // then x-ghosts, y-ghosts, z-ghosts; -1 when no such cell exists on this block.
  inline long cell_id(int i, int j, int k) const {
    const int nx = b.n[0], ny = b.n[1], nz = b.n[2];
    const bool ix = (i >= 0 && i < nx), iy = (j >= 0 && j < ny), iz = (k >= 0 && k < nz);
    if (ix && iy && iz) return ((long)i * ny + j) * nz + k;
    if (!ix && iy && iz) {
      if (i == -1 && b.glo[0]) return base[0] + ((long)0 * ny + j) * nz + k;
      if (i == nx && b.ghi[0]) return base[0] + ((long)b.glo[0] * ny + j) * nz + k;
      return -1;
    }
    if (ix && !iy && iz) {
      if (j == -1 && b.glo[1]) return base[1] + ((long)i * ng[1] + 0) * nz + k;
      if (j == ny && b.ghi[1]) return base[1] + ((long)i * ng[1] + b.glo[1]) * nz + k;
      return -1;
    }
    if (ix && iy && !iz) {
      if (k == -1 && b.glo[2]) return base[2] + ((long)i * ny + j) * ng[2] + 0;
      if (k == nz && b.ghi[2]) return base[2] + ((long)i * ny + j) * ng[2] + b.glo[2];
      return -1;
    }
    return -1;
  }
  // inverse of cell_id
  inline void cell_ijk(long id, int &i, int &j, int &k) const {
    const int nx = b.n[0], ny = b.n[1], nz = b.n[2];
    if (id < nowned) {
      k = (int)(id % nz);
      long r = id / nz;
      j = (int)(r % ny);
      i = (int)(r / ny);
      return;
    }
    if (id < base[1]) {
      long r = id - base[0];
      k = (int)(r % nz);
      r /= nz;
      j = (int)(r % ny);
      int xi = (int)(r / ny);
      i = (xi == 0 && b.glo[0]) ? -1 : nx;
      return;
    }
    if (id < base[2]) {
      long r = id - base[1];
      k = (int)(r % nz);
      r /= nz;
      int yj = (int)(r % ng[1]);
      i = (int)(r / ng[1]);
      j = (yj == 0 && b.glo[1]) ? -1 : ny;
      return;
    }
    long r = id - base[2];
    int zk = (int)(r % ng[2]);
    r /= ng[2];
    j = (int)(r % ny);
    i = (int)(r / ny);
    k = (zk == 0 && b.glo[2]) ? -1 : nz;
  }
  inline int global_id(int i, int j, int k) const {  // Parallel3DMesh.h:462-464
    return (b.off[0] + i) * (b.gn[1] * b.gn[2]) + (b.off[1] + j) * b.gn[2] + (b.off[2] + k);
  }

  // Face.C:37-98 for local face `f` of cell (i,j,k)
  void face_geometry(int i, int j, int k, int f, double *coords, double *a, double *t, double *bn) const {
    double n[4][3];
    for (int q = 0; q < 4; ++q) {
      const int *o = kNodeOff[kFaceNodes[f][q]];
      node(i + o[0], j + o[1], k + o[2], n[q]);
    }
    coords[0] = coords[1] = coords[2] = 0.0;
    for (int q = 0; q < 4; ++q) {
      coords[0] += n[q][0];
      coords[1] += n[q][1];
      coords[2] += n[q][2];
    }
    const double s = 1.0 / 4;
    coords[0] *= s, coords[1] *= s, coords[2] *= s;
    const double v1[3] = {n[1][0] - n[0][0], n[1][1] - n[0][1], n[1][2] - n[0][2]};
    const double v2[3] = {n[2][0] - n[0][0], n[2][1] - n[0][1], n[2][2] - n[0][2]};
    const double v3[3] = {n[3][0] - n[0][0], n[3][1] - n[0][1], n[3][2] - n[0][2]};
    double n1[3], n2[3];
    // MathTools.h:48-53 Vec3Cross
    n1[0] = v1[1] * v2[2] - v2[1] * v1[2];
    n1[1] = -v1[0] * v2[2] + v2[0] * v1[2];
    n1[2] = v1[0] * v2[1] - v2[0] * v1[1];
    n2[0] = v2[1] * v3[2] - v3[1] * v2[2];
    n2[1] = -v2[0] * v3[2] + v3[0] * v2[2];
    n2[2] = v2[0] * v3[1] - v3[0] * v2[1];
    a[0] = 0.5 * (n1[0] + n2[0]);
    a[1] = 0.5 * (n1[1] + n2[1]);
    a[2] = 0.5 * (n1[2] + n2[2]);
    // tangent: Face.C:81-92 (std::max_element returns the FIRST largest)
    const double ab[3] = {std::abs(a[0]), std::abs(a[1]), std::abs(a[2])};
    int i1 = 0;
    if (ab[1] > ab[i1]) i1 = 1;
    if (ab[2] > ab[i1]) i1 = 2;
    int i2 = i1 + 1, i3 = i1 + 2;
    i2 = (i2 > 2) ? i2 - 3 : i2;
    i3 = (i3 > 2) ? i3 - 3 : i3;
    const double denom = std::sqrt(a[i1] * a[i1] + a[i3] * a[i3]);
    t[i2] = 0.0;
    t[i1] = a[i3] / denom;
    t[i3] = -a[i1] / denom;
    bn[0] = a[1] * t[2] - t[1] * a[2];
    bn[1] = -a[0] * t[2] + t[0] * a[2];
    bn[2] = a[0] * t[1] - t[0] * a[1];
  }

  // MeshProcessor.C:130-171 + ElementTopoHexa8.C:52-148
  void cell_geometry(int i, int j, int k, double *centroid, double *volume) const {
    double ex[8], ey[8], ez[8];
    for (int q = 0; q < 8; ++q) {
      double c[3];
      node(i + kNodeOff[q][0], j + kNodeOff[q][1], k + kNodeOff[q][2], c);
      ex[q] = c[0], ey[q] = c[1], ez[q] = c[2];
    }
    double sx = 0, sy = 0, sz = 0;
    for (int q = 0; q < 8; ++q) {
      sx += ex[q];
      sy += ey[q];
      sz += ez[q];
    }
    centroid[0] = sx / 8;
    centroid[1] = sy / 8;
    centroid[2] = sz / 8;
    double vol = 0.0;
    for (int g = 0; g < 8; ++g) {
      const double xi = kXi[g], eta = kEta[g], zeta = kZeta[g];
      double dxi[8], deta[8], dzeta[8];
      dxi[0] = -0.125 * (1.0 - eta) * (1.0 - zeta);
      dxi[1] = 0.125 * (1.0 - eta) * (1.0 - zeta);
      dxi[2] = 0.125 * (1.0 + eta) * (1.0 - zeta);
      dxi[3] = -0.125 * (1.0 + eta) * (1.0 - zeta);
      dxi[4] = -0.125 * (1.0 - eta) * (1.0 + zeta);
      dxi[5] = 0.125 * (1.0 - eta) * (1.0 + zeta);
      dxi[6] = 0.125 * (1.0 + eta) * (1.0 + zeta);
      dxi[7] = -0.125 * (1.0 + eta) * (1.0 + zeta);
      deta[0] = -0.125 * (1.0 - xi) * (1.0 - zeta);
      deta[1] = -0.125 * (1.0 + xi) * (1.0 - zeta);
      deta[2] = 0.125 * (1.0 + xi) * (1.0 - zeta);
      deta[3] = 0.125 * (1.0 - xi) * (1.0 - zeta);
      deta[4] = -0.125 * (1.0 - xi) * (1.0 + zeta);
      deta[5] = -0.125 * (1.0 + xi) * (1.0 + zeta);
      deta[6] = 0.125 * (1.0 + xi) * (1.0 + zeta);
      deta[7] = 0.125 * (1.0 - xi) * (1.0 + zeta);
      dzeta[0] = -0.125 * (1.0 - xi) * (1.0 - eta);
      dzeta[1] = -0.125 * (1.0 + xi) * (1.0 - eta);
      dzeta[2] = -0.125 * (1.0 + xi) * (1.0 + eta);
      dzeta[3] = -0.125 * (1.0 - xi) * (1.0 + eta);
      dzeta[4] = 0.125 * (1.0 - xi) * (1.0 - eta);
      dzeta[5] = 0.125 * (1.0 + xi) * (1.0 - eta);
      dzeta[6] = 0.125 * (1.0 + xi) * (1.0 + eta);
      dzeta[7] = 0.125 * (1.0 - xi) * (1.0 + eta);
      double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int q = 0; q < 8; ++q) J[0] += dxi[q] * ex[q];
      for (int q = 0; q < 8; ++q) J[1] += dxi[q] * ey[q];
      for (int q = 0; q < 8; ++q) J[2] += dxi[q] * ez[q];
      for (int q = 0; q < 8; ++q) J[3] += deta[q] * ex[q];
      for (int q = 0; q < 8; ++q) J[4] += deta[q] * ey[q];
      for (int q = 0; q < 8; ++q) J[5] += deta[q] * ez[q];
      for (int q = 0; q < 8; ++q) J[6] += dzeta[q] * ex[q];
      for (int q = 0; q < 8; ++q) J[7] += dzeta[q] * ey[q];
      for (int q = 0; q < 8; ++q) J[8] += dzeta[q] * ez[q];
      const double detJ = J[0] * (J[4] * J[8] - J[5] * J[7]) + J[1] * (J[5] * J[6] - J[3] * J[8]) +
                          J[2] * (J[3] * J[7] - J[4] * J[6]);
      vol += detJ;
    }
    *volume = vol;
  }
};

struct FaceArrays {
  std::vector<double> coords, normal, tangent, binormal;
  std::vector<int> conn, slot;
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
  std::vector<double> cell_xyz, cell_vol;
  std::vector<int> global_ids;
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
    g.lx = opt->lx, g.ly = opt->ly, g.lz = opt->lz;
    const double PI = 3.14159265;  // Parallel3DMesh.h:473
    g.tan_ramp = std::tan(opt->angle * PI / 180.0);
    g.build_tables();
    const int nx = b.n[0], ny = b.n[1], nz = b.n[2];
    g.nowned = (long)nx * ny * nz;
    for (int d = 0; d < 3; ++d) g.ng[d] = b.glo[d] + b.ghi[d];
    g.base[0] = g.nowned;
    g.base[1] = g.base[0] + (long)g.ng[0] * ny * nz;
    g.base[2] = g.base[1] + (long)g.ng[1] * nx * nz;
    const long ncells = g.base[2] + (long)g.ng[2] * nx * ny;
    g.nghost = ncells - g.nowned;
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
    std::vector<long> first((size_t)ncells + 1, 0);
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
