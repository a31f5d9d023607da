// Structured-block geometry and numbering shared by the host mesh generator (host_mesh.cpp), the structured layout
// builder (layout.cpp) and the device-side geometry kernels (geom_kernels.cu): the reference's Parallel3DMesh
// block decomposition and cell numbering, Face.C's face geometry and ElementTopoHexa8's cell volume, as plain
// functions of (i, j, k) over node-coordinate tables.  Everything is evaluated in the reference's order; every
// translation unit that includes this header is compiled without FMA contraction (-ffp-contract=off / -fmad=false),
// so host and device produce the same bits.
#pragma once
#include <cmath>
#include <vector>

#ifdef __CUDACC__
#define MA_HD __host__ __device__
#else
#define MA_HD
#endif

namespace ma {

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

inline bool arrange(Block &b, int gnx, int gny, int gnz, int rank, int nprocs) {
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

// local face f of a hex: direction of the neighbour across it and the slot the face has in that neighbour
// 0: -y, 1: +x, 2: +y, 3: -x, 4: -z, 5: +z   (MeshProcessor.C:44; slot s of a cell is local face s)
MA_HD inline void face_dir(int f, int &di, int &dj, int &dk) {
  di = (f == 1) - (f == 3);
  dj = (f == 2) - (f == 0);
  dk = (f == 5) - (f == 4);
}
MA_HD inline int opposite_face(int f) { return f < 4 ? (f + 2) & 3 : 9 - f; }

// Plain-data view of one block of the structured mesh: usable on the host and on the device.  xs / zs are indexed by
// global node index + 1 (-1 .. gn + 1), ys by (global i + 1) * ysj + (global j + 1) (Parallel3DMesh.h:466-487).
struct GridGen {
  Block b;
  const double *xs = nullptr, *ys = nullptr, *zs = nullptr;
  int ysj = 0;
  long nowned = 0, nghost = 0;
  long base[3] = {0, 0, 0};  // first local id of the x-, y-, z-ghost groups
  int ng[3] = {0, 0, 0};     // ghost layers per direction (0..2)

  void set_block_counts() {
    const int nx = b.n[0], ny = b.n[1], nz = b.n[2];
    nowned = (long)nx * ny * nz;
    for (int d = 0; d < 3; ++d) ng[d] = b.glo[d] + b.ghi[d];
    base[0] = nowned;
    base[1] = base[0] + (long)ng[0] * ny * nz;
    base[2] = base[1] + (long)ng[1] * nx * nz;
    nghost = base[2] + (long)ng[2] * nx * ny - nowned;
  }
  MA_HD inline void node(int i, int j, int k, double *c) const {  // local node index -> coordinate
    const int gi = b.off[0] + i, gj = b.off[1] + j, gk = b.off[2] + k;
    c[0] = xs[gi + 1];
    c[1] = ys[(size_t)(gi + 1) * ysj + (gj + 1)];
    c[2] = zs[gk + 1];
  }

  // local cell id in the reference's numbering (Parallel3DMesh.C:80-174): owned cells k-fastest,
  // then x-ghosts, y-ghosts, z-ghosts; -1 when no such cell exists on this block.
  MA_HD inline long cell_id(int i, int j, int k) const {
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
  MA_HD inline void cell_ijk(long id, int &i, int &j, int &k) const {
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
  MA_HD inline int global_id(int i, int j, int k) const {  // Parallel3DMesh.h:462-464
    return (b.off[0] + i) * (b.gn[1] * b.gn[2]) + (b.off[1] + j) * b.gn[2] + (b.off[2] + k);
  }

  // Face.C:37-98 for local face `f` of cell (i,j,k)
  MA_HD inline void face_geometry(int i, int j, int k, int f, double *coords, double *a, double *t, double *bn) const {
    // Hex8 local face -> local nodes (MeshProcessor.C:44); node n of cell (i,j,k) is (i+di, j+dj, k+dk)
    // (Parallel3DMesh.C:83-95)
    const int kFaceNodes[6][4] = {{0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {0, 3, 2, 1}, {4, 5, 6, 7}};
    const int kNodeOff[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
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
    const double ab[3] = {fabs(a[0]), fabs(a[1]), fabs(a[2])};
    int i1 = 0;
    if (ab[1] > ab[i1]) i1 = 1;
    if (ab[2] > ab[i1]) i1 = 2;
    int i2 = i1 + 1, i3 = i1 + 2;
    i2 = (i2 > 2) ? i2 - 3 : i2;
    i3 = (i3 > 2) ? i3 - 3 : i3;
    const double denom = sqrt(a[i1] * a[i1] + a[i3] * a[i3]);
    t[i2] = 0.0;
    t[i1] = a[i3] / denom;
    t[i3] = -a[i1] / denom;
    bn[0] = a[1] * t[2] - t[1] * a[2];
    bn[1] = -a[0] * t[2] + t[0] * a[2];
    bn[2] = a[0] * t[1] - t[0] * a[1];
  }

  // MeshProcessor.C:130-171 + ElementTopoHexa8.C:52-148
  MA_HD inline void cell_geometry(int i, int j, int k, double *centroid, double *volume) const {
    const int kNodeOff[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    // 2x2x2 Gauss points (ElementTopoHexa8.h:35-43)
    const double kG = 0.577350269189626;
    const double kXi[8] = {-kG, kG, kG, -kG, -kG, kG, kG, -kG};
    const double kEta[8] = {-kG, -kG, kG, kG, -kG, -kG, kG, kG};
    const double kZeta[8] = {-kG, -kG, -kG, -kG, kG, kG, kG, kG};
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

// Host storage of the node-coordinate tables (Parallel3DMesh.h:466-487 — note the two different expressions for y
// either side of lx/2).  The tables cover the global index range -1 .. gn + 1, so they are the same on every rank.
struct GridTables {
  std::vector<double> xs, ys, zs;
  void build(GridGen &g, double lx, double ly, double lz, double tan_ramp) {
    const int NX = g.b.gn[0], NY = g.b.gn[1], NZ = g.b.gn[2];
    xs.resize(NX + 3);
    zs.resize(NZ + 3);
    for (int gi = -1; gi <= NX + 1; ++gi) xs[gi + 1] = (double)gi / (NX)*lx;
    for (int gk = -1; gk <= NZ + 1; ++gk) zs[gk + 1] = (double)gk / (NZ)*lz;
    const int ysj = NY + 3;
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
    g.xs = xs.data(), g.ys = ys.data(), g.zs = zs.data();
    g.ysj = ysj;
  }
};

}  // namespace ma
