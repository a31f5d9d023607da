// See layout.h.  Everything here is one-time host setup (the analogue of the reference's
// copy_faces / copy_cell_data, Faces.h:88-145, Cells.h:126-146); it is O(cells) and OpenMP-parallel.
#include "layout.h"

#include <algorithm>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#include <parallel/algorithm>
#endif

#include "host_common.h"

namespace ma {

namespace {

struct FaceSrc {  // where global face g lives in the caller's arrays
  const ma_faces *f;
  int index;
  int bc_type;  // -1: internal
};

inline int round_up(long v, int m) { return (int)(((v + m - 1) / m) * m); }

// 3 x 21-bit Morton (Z-order) key: consecutive tiles are spatial neighbours in all three directions, so the
// cells a tile reads across its cut faces were touched by a recently scheduled CTA and are still in L2
inline uint64_t spread3(uint64_t v) {
  v &= 0x1fffffULL;
  v = (v | v << 32) & 0x1f00000000ffffULL;
  v = (v | v << 16) & 0x1f0000ff0000ffULL;
  v = (v | v << 8) & 0x100f00f00f00f00fULL;
  v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
  v = (v | v << 2) & 0x1249249249249249ULL;
  return v;
}

// Bank-conflict-aware order of one group of tile faces (FAST staged flux kernel).  Work item q of the group is
// evaluated by lane (lane0 + q) of the CTA; a 64-bit shared-memory load is served half-warp by half-warp, one
// wavefront when the 16 lanes address 16 different double-words modulo 16.  Every item reads the staged records of
// up to two cells (positions a, b; -1 = none) in load streams `sa`, `sb` (lanes of different streams execute
// different instructions: interior / boundary / cut-left / cut-right).  Greedy: walk the lanes in order and give
// each the earliest unplaced item (within a window) whose positions collide with nothing already in its half-warp.
// The slot-ordered gather makes results independent of the face order (tests: test_result_independent_of_tiling).
struct PackItem {
  int a, b;   // staged-cell positions read by this face (b < 0: only a)
  int sa, sb; // stream of each read, 0..3
};
inline void pack_conflict_free(const std::vector<PackItem> &items, int lane0, std::vector<int> &order) {
  const int n = (int)items.size();
  order.clear();
  order.reserve(n);
  // lanes of one kind (code path) stay together: a warp that holds two kinds executes both paths
  std::vector<int> seg;
  uint16_t mask[4] = {0, 0, 0, 0};
  const int window = 128;
  for (int kind = 0; kind < 4; ++kind) {
    seg.clear();
    for (int j = 0; j < n; ++j)
      if (items[j].sa == kind) seg.push_back(j);
    const int m = (int)seg.size();
    std::vector<char> used(m, 0);
    int first = 0;
    for (int q = 0; q < m; ++q) {
      if (((lane0 + (int)order.size()) & 15) == 0) mask[0] = mask[1] = mask[2] = mask[3] = 0;
      while (first < m && used[first]) ++first;
      int pick = first, seen = 0;
      for (int j = first; j < m && seen < window; ++j) {
        if (used[j]) continue;
        ++seen;
        const PackItem &it = items[seg[j]];
        if (mask[it.sa] >> (it.a & 15) & 1) continue;
        if (it.b >= 0 && (mask[it.sb] >> (it.b & 15) & 1)) continue;
        pick = j;
        break;
      }
      const PackItem &it = items[seg[pick]];
      mask[it.sa] |= (uint16_t)(1u << (it.a & 15));
      if (it.b >= 0) mask[it.sb] |= (uint16_t)(1u << (it.b & 15));
      used[pick] = 1;
      order.push_back(seg[pick]);
    }
  }
}

inline uint64_t morton3(long x, long y, long z) { return spread3((uint64_t)x) << 2 | spread3((uint64_t)y) << 1 | spread3((uint64_t)z); }


// What the builder asks about (owned cell c, slot s)
struct SlotInfo {
  int other;       // cell across the face (caller's numbering): -1 boundary face, >= n_owned ghost
  int bc_type;     // ma_bc_type of a boundary face, -1 otherwise
  int side;        // 0: c is elem1 (the normal points out of c), 1: c is elem2
  int other_slot;  // slot of the face in the cell across (interior faces)
};

// ---- a mesh handed over as reference-format arrays (ma_mesh) ---------------------------------------------
struct ArrayAccess {
  const ma_mesh &mesh;
  int n_owned = 0, n_ghost = 0;
  long n_cells = 0, n_int = 0;
  std::vector<long> set_base;
  std::vector<uint32_t> cf;  // (cell, slot) -> global_face * 2 + side; internal faces first, then the boundary sets
  double lo[3], h[3];
  long nbins[3];
  static constexpr bool kStructured = false;

  explicit ArrayAccess(const ma_mesh &m) : mesh(m) {}

  int prepare() {
    n_owned = mesh.num_owned_cells, n_ghost = mesh.num_ghosts;
    n_cells = (long)n_owned + n_ghost;
    if (n_owned <= 0 || n_ghost < 0) return ma_set_error(MA_ERR_INVALID, "mesh: num_owned_cells must be > 0");
    if (!mesh.cell_coordinates || !mesh.cell_volumes)
      return ma_set_error(MA_ERR_INVALID, "mesh: cell_coordinates / cell_volumes are NULL");
    if (mesh.num_boundary_sets < 0 || mesh.num_boundary_sets > MA_MAX_BC_SETS)
      return ma_set_error(MA_ERR_INVALID, "mesh: num_boundary_sets out of range");
    auto faces_ok = [](const ma_faces &f) {
      return f.nfaces == 0 || (f.nfaces > 0 && f.coordinates && f.face_normal && f.face_tangent && f.face_binormal &&
                               f.face_cell_conn && f.cell_flux_index);
    };
    if (!faces_ok(mesh.internal_faces)) return ma_set_error(MA_ERR_INVALID, "mesh: internal_faces has NULL arrays");
    for (int b = 0; b < mesh.num_boundary_sets; ++b) {
      if (!faces_ok(mesh.boundary_faces[b])) return ma_set_error(MA_ERR_INVALID, "mesh: boundary set has NULL arrays");
      if (mesh.boundary_type[b] < 0 || mesh.boundary_type[b] > 3)
        return ma_set_error(MA_ERR_INVALID, "mesh: unknown boundary_type");
    }
    if (n_ghost > 0 && (mesh.num_ranks < 2 || !mesh.send_count || !mesh.recv_count || !mesh.send_local_ids ||
                        !mesh.recv_local_ids))
      return ma_set_error(MA_ERR_INVALID, "mesh: ghosts present but exchange lists are missing");

    // cell -> face table over owned cells
    n_int = mesh.internal_faces.nfaces;
    set_base.assign(mesh.num_boundary_sets + 1, n_int);
    for (int b = 0; b < mesh.num_boundary_sets; ++b) set_base[b + 1] = set_base[b] + mesh.boundary_faces[b].nfaces;
    const long n_faces_all = set_base[mesh.num_boundary_sets];
    if (n_faces_all >= (1L << 31)) return ma_set_error(MA_ERR_INVALID, "mesh: more than 2^31 faces");
    const uint32_t kNone = 0xFFFFFFFFu;
    cf.assign((size_t)n_owned * 6, kNone);
    int bad = 0;
    {
      const int *conn = mesh.internal_faces.face_cell_conn, *slot = mesh.internal_faces.cell_flux_index;
      const long ncl = n_cells;
      const int no = n_owned;
#pragma omp parallel for schedule(static) reduction(+ : bad)
      for (long f = 0; f < n_int; ++f) {
        for (int side = 0; side < 2; ++side) {
          const int c = conn[2 * f + side], s = slot[2 * f + side];
          if (c < 0 || c >= ncl || s < 0 || s > 5) {
            ++bad;
            continue;
          }
          if (c < no) cf[(size_t)c * 6 + s] = (uint32_t)(f * 2 + side);
        }
      }
      for (int b = 0; b < mesh.num_boundary_sets; ++b) {
        const ma_faces &F = mesh.boundary_faces[b];
        const long base = set_base[b];
#pragma omp parallel for schedule(static) reduction(+ : bad)
        for (long f = 0; f < F.nfaces; ++f) {
          const int c = F.face_cell_conn[2 * f], s = F.cell_flux_index[2 * f];
          if (c < 0 || c >= no || s < 0 || s > 5) {
            ++bad;
            continue;
          }
          cf[(size_t)c * 6 + s] = (uint32_t)((base + f) * 2);
        }
      }
    }
    if (bad) return ma_set_error(MA_ERR_INVALID, "mesh: face_cell_conn / cell_flux_index out of range");
    {
      long missing = 0;
#pragma omp parallel for schedule(static) reduction(+ : missing)
      for (long i = 0; i < (long)n_owned * 6; ++i) missing += (cf[i] == kNone);
      if (missing)
        return ma_set_error(MA_ERR_INVALID, "mesh: " + std::to_string(missing) +
                                                " (cell, slot) pairs of owned cells have no face (hex cells need 6)");
    }
    return MA_OK;
  }

  struct FaceSrc {  // where a face lives in the caller's arrays
    const ma_faces *f;
    long index;
    int bc_type;  // -1: internal
  };
  FaceSrc face_src(uint32_t ref) const {
    const long g = ref >> 1;
    FaceSrc s;
    if (g < n_int) {
      s.f = &mesh.internal_faces, s.index = g, s.bc_type = -1;
    } else {
      int b = 0;
      while (g >= set_base[b + 1]) ++b;
      s.f = &mesh.boundary_faces[b], s.index = g - set_base[b], s.bc_type = mesh.boundary_type[b];
    }
    return s;
  }
  SlotInfo info(int c, int s) const {
    const uint32_t ref = cf[(size_t)c * 6 + s];
    const long g = ref >> 1;
    SlotInfo r;
    r.side = (int)(ref & 1);
    if (g >= n_int) {
      r.other = -1, r.other_slot = -1;
      r.bc_type = face_src(ref).bc_type;
    } else {
      r.other = mesh.internal_faces.face_cell_conn[2 * g + (1 - r.side)];
      r.other_slot = mesh.internal_faces.cell_flux_index[2 * g + (1 - r.side)];
      r.bc_type = -1;
    }
    return r;
  }
  // area vector, tangent, binormal, centroid of the face in slot s of owned cell c
  void face_geometry(int c, int s, double *n, double *t, double *b, double *x) const {
    const FaceSrc src = face_src(cf[(size_t)c * 6 + s]);
    const size_t fi = (size_t)src.index;
    for (int d = 0; d < 3; ++d) {
      n[d] = src.f->face_normal[3 * fi + d];
      t[d] = src.f->face_tangent[3 * fi + d];
      b[d] = src.f->face_binormal[3 * fi + d];
      x[d] = src.f->coordinates[3 * fi + d];
    }
  }
  uint32_t face_code(int, int) const { return 0; }
  uint64_t tile_key(int, const int *, int, int) const { return 0; }
  void cell_geometry(long c, double *xyz, double *vol) const {
    for (int d = 0; d < 3; ++d) xyz[d] = mesh.cell_coordinates[3 * c + d];
    *vol = mesh.cell_volumes[c];
  }

  // spatial binning of owned cells.  The mean centroid spacing along each axis is taken over internal faces whose
  // cell-to-cell vector is dominated by that axis; for a structured block this recovers (i,j,k) exactly, for a
  // general hex mesh it only has to give compact tiles.
  void prepare_bins() {
    const double *xc = mesh.cell_coordinates;
    double hi[3] = {-1e300, -1e300, -1e300};
    lo[0] = lo[1] = lo[2] = 1e300;
    for (long c = 0; c < n_owned; ++c)
      for (int d = 0; d < 3; ++d) {
        lo[d] = std::min(lo[d], xc[3 * c + d]);
        hi[d] = std::max(hi[d], xc[3 * c + d]);
      }
    double hsum[3] = {0, 0, 0};
    long hcnt[3] = {0, 0, 0};
    {
      const int *conn = mesh.internal_faces.face_cell_conn;
      double s0 = 0, s1 = 0, s2 = 0;
      long c0 = 0, c1 = 0, c2 = 0;
      const int no = n_owned;
#pragma omp parallel for schedule(static) reduction(+ : s0, s1, s2, c0, c1, c2)
      for (long f = 0; f < n_int; ++f) {
        const int l = conn[2 * f], r = conn[2 * f + 1];
        if (l >= no || r >= no) continue;
        const double d0 = std::fabs(xc[3 * (long)r] - xc[3 * (long)l]);
        const double d1 = std::fabs(xc[3 * (long)r + 1] - xc[3 * (long)l + 1]);
        const double d2 = std::fabs(xc[3 * (long)r + 2] - xc[3 * (long)l + 2]);
        if (d0 >= d1 && d0 >= d2) {
          s0 += d0, ++c0;
        } else if (d1 >= d2) {
          s1 += d1, ++c1;
        } else {
          s2 += d2, ++c2;
        }
      }
      hsum[0] = s0, hsum[1] = s1, hsum[2] = s2;
      hcnt[0] = c0, hcnt[1] = c1, hcnt[2] = c2;
    }
    for (int d = 0; d < 3; ++d) {
      h[d] = hcnt[d] ? hsum[d] / (double)hcnt[d] : 0.0;
      if (!(h[d] > 0.0) || !((hi[d] - lo[d]) / h[d] < 1e9)) h[d] = (hi[d] - lo[d]) + 1.0;  // one bin
      nbins[d] = (long)std::floor((hi[d] - lo[d]) / h[d] + 0.5) + 1;
    }
  }
  void bin(long c, long q[3]) const {
    const double *xc = mesh.cell_coordinates;
    for (int d = 0; d < 3; ++d) {
      q[d] = (long)std::floor((xc[3 * c + d] - lo[d]) / h[d] + 0.5);
      q[d] = std::max(0L, std::min(q[d], nbins[d] - 1));
    }
  }
  int num_ranks() const { return mesh.num_ranks; }
  int my_rank() const { return mesh.my_rank; }
  // exchange lists of peer p: counts, and the i-th send / recv cell (caller's numbering) from running offsets
  int send_count(int p) const { return mesh.send_count[p]; }
  int recv_count(int p) const { return mesh.recv_count[p]; }
  int send_id(long off) const { return mesh.send_local_ids[off]; }
  int recv_id(long off) const { return mesh.recv_local_ids[off]; }
};

// ---- one block of the in-code structured mesh, never materialised (mesh_geom.h) ------------------------------
struct StructuredAccess {
  GridGen g;
  GridTables tables;
  int bc_of_face[6];  // ma_bc_type of the domain side behind local face f (Parallel3DMesh.h:382-396)
  int n_owned = 0, n_ghost = 0;
  long n_cells = 0;
  long nbins[3];
  std::vector<int> send_ids, recv_ids, send_counts, recv_counts;
  static constexpr bool kStructured = true;

  int prepare(const ma_options &opt, int rank, int num_ranks) {
    if (opt.nx <= 0 || opt.ny <= 0 || opt.nz <= 0)
      return ma_set_error(MA_ERR_INVALID, "structured mesh: nx, ny, nz must be positive");
    if (num_ranks < 1 || rank < 0 || rank >= num_ranks) return ma_set_error(MA_ERR_INVALID, "structured mesh: bad rank / num_ranks");
    if (!arrange(g.b, opt.nx, opt.ny, opt.nz, rank, num_ranks))
      return ma_set_error(MA_ERR_INVALID, "MPI number of ranks must be a power of 2.");  // Parallel3DMesh.C:262-265
    for (int d = 0; d < 3; ++d)
      if (g.b.n[d] < 1) return ma_set_error(MA_ERR_INVALID, "structured mesh: more blocks than cells in a direction");
    const double PI = 3.14159265;  // Parallel3DMesh.h:473
    tables.build(g, opt.lx, opt.ly, opt.lz, std::tan(opt.angle * PI / 180.0));
    g.set_block_counts();
    if ((g.nowned + g.nghost) * 3 > 2000000000L)
      return ma_set_error(MA_ERR_INVALID, "structured mesh: more than 2^31 faces on one block");
    n_owned = (int)g.nowned, n_ghost = (int)g.nghost, n_cells = g.nowned + g.nghost;
    for (int d = 0; d < 3; ++d) nbins[d] = g.b.n[d];
    const int pt = opt.problem_type;
    bc_of_face[0] = (pt == 1) ? MA_BC_NOSLIP : MA_BC_TANGENT;       // bottom (-y)
    bc_of_face[2] = (pt == 1) ? MA_BC_EXTRAPOLATE : MA_BC_TANGENT;  // top (+y)
    bc_of_face[4] = MA_BC_TANGENT;                                  // front (-z)
    bc_of_face[5] = MA_BC_TANGENT;                                  // back (+z)
    bc_of_face[1] = MA_BC_EXTRAPOLATE;                              // right (+x)
    bc_of_face[3] = (pt == 0) ? MA_BC_EXTRAPOLATE : MA_BC_INFLOW;   // left (-x)
    // ghost exchange lists: per neighbour rank (ascending), ordered by global id (host_mesh.cpp)
    send_counts.assign(num_ranks, 0);
    recv_counts.assign(num_ranks, 0);
    if (num_ranks > 1) {
      const Block &b = g.b;
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
        send_counts[nb.rank] = (int)cnt;
        recv_counts[nb.rank] = (int)cnt;
        for (long w = 0; w < cnt; ++w) {
          int idx[3];
          idx[a1] = (int)(w / b.n[a2]);
          idx[a2] = (int)(w % b.n[a2]);
          idx[d] = nb.hi ? b.n[d] - 1 : 0;
          send_ids.push_back((int)g.cell_id(idx[0], idx[1], idx[2]));
          idx[d] = nb.hi ? b.n[d] : -1;
          recv_ids.push_back((int)g.cell_id(idx[0], idx[1], idx[2]));
        }
      }
    }
    num_ranks_ = num_ranks, my_rank_ = rank;
    return MA_OK;
  }
  int num_ranks_ = 1, my_rank_ = 0;

  SlotInfo info(int c, int s) const {
    int i, j, k, di, dj, dk;
    g.cell_ijk(c, i, j, k);
    face_dir(s, di, dj, dk);
    const long nb = g.cell_id(i + di, j + dj, k + dk);
    SlotInfo r;
    if (nb < 0) {
      r.other = -1, r.other_slot = -1, r.side = 0, r.bc_type = bc_of_face[s];
    } else {
      r.other = (int)nb, r.other_slot = opposite_face(s), r.bc_type = -1;
      r.side = nb > c ? 0 : 1;  // a face is created by the lower-numbered of its two cells (MeshProcessor.C:54-120)
    }
    return r;
  }
  // the cell that created the face (elem1) and the local face it created it as: Face.C computes the geometry there
  void elem1(int c, int s, int &i, int &j, int &k, int &f) const {
    int di, dj, dk;
    g.cell_ijk(c, i, j, k);
    face_dir(s, di, dj, dk);
    const long nb = g.cell_id(i + di, j + dj, k + dk);
    f = s;
    if (nb >= 0 && nb < c) {
      i += di, j += dj, k += dk;
      f = opposite_face(s);
    }
  }
  void face_geometry(int c, int s, double *n, double *t, double *b, double *x) const {
    int i, j, k, f;
    elem1(c, s, i, j, k, f);
    g.face_geometry(i, j, k, f, x, n, t, b);
  }
  // (elem1 cell in the block-local (i+1, j+1, k+1) lattice) * 8 + elem1 local face: what the device-side geometry
  // kernel needs to recompute the face (geom_kernels.cu)
  uint32_t face_code(int c, int s) const {
    int i, j, k, f;
    elem1(c, s, i, j, k, f);
    const long lat = ((long)(i + 1) * (g.b.n[1] + 2) + (j + 1)) * (g.b.n[2] + 2) + (k + 1);
    return (uint32_t)(lat * 8 + f);
  }
  void cell_geometry(long c, double *xyz, double *vol) const {
    int i, j, k;
    g.cell_ijk(c, i, j, k);
    g.cell_geometry(i, j, k, xyz, vol);
  }
  void prepare_bins() {}
  void bin(long c, long q[3]) const {
    int i, j, k;
    g.cell_ijk(c, i, j, k);
    q[0] = i, q[1] = j, q[2] = k;
  }
  // Everything the face list of a tile depends on, for a tile that is a brick of the block: its extents, what lies
  // behind each of its six sides (another tile of the block / a ghost layer / the domain boundary) and the parity of
  // its first cell (the staged positions are shifted by it).  Tiles with equal keys have the same tile-local face
  // order, so the builder computes it once per key (a few dozen keys per block).  0 = not a brick: no sharing.
  uint64_t tile_key(int first_cell, const int td[3], int cell_count, int parity) const {
    int ijk[3];
    g.cell_ijk(first_cell, ijk[0], ijk[1], ijk[2]);
    uint64_t key = 1;
    long cells = 1;
    for (int d = 0; d < 3; ++d) {
      const int origin = ijk[d] / td[d] * td[d];
      const int ext = std::min(td[d], g.b.n[d] - origin);
      cells *= ext;
      const int lo = origin > 0 ? 0 : (g.b.glo[d] ? 1 : 2);
      const int hi = origin + ext < g.b.n[d] ? 0 : (g.b.ghi[d] ? 1 : 2);
      key = key << 16 | (uint64_t)(ext & 0xfff) << 4 | (uint64_t)(lo << 2 | hi);
    }
    if (cells != cell_count) return 0;
    return key << 1 | (uint64_t)(parity & 1);
  }
  int num_ranks() const { return num_ranks_; }
  int my_rank() const { return my_rank_; }
  int send_count(int p) const { return send_counts[p]; }
  int recv_count(int p) const { return recv_counts[p]; }
  int send_id(long off) const { return send_ids[off]; }
  int recv_id(long off) const { return recv_ids[off]; }
};

// The tile-local face order of one tile: (tile-local cell, slot) of every tile face; group 0 closed / boundary, 1 cut and
// evaluated here, 2 cut and imported (shared cut faces).
struct TileOrder {
  std::vector<uint16_t> lc[3];
  std::vector<uint8_t> slot[3];
  int dummy_group = -1;  // the LAST entry of this group repeats an evaluated face (pads n_eval to even): it is
                         // evaluated, but no cell's slot refers to it
};
// Inside each group the faces are listed by slot (direction), then by cell — consecutive faces read consecutive own
// cells and consecutive neighbours — and, for the staged FAST kernels (`pack`), that list is re-packed half-warp by
// half-warp so that the 16 lanes of a shared-memory wavefront read 16 different banks (pack_conflict_free).  `by_cell`
// (STRICT, MINIAERO_FACE_ORDER=cell) keeps (cell, slot) order (the gather kernels' L1 locality).
// The tile is seen through three callbacks over tile-local cell indices, so the same code serves the builder that
// works on the renumbered global arrays and the one that derives a brick's pattern from (i, j, k) alone:
//   emits(lc, s)  does cell lc list the face behind its slot s (the in-tile cell with the larger local index does)
//   kind(lc, s)   0 = closed or boundary face, 1 = cut face evaluated by this tile, 2 = cut face imported
//   pinfo(lc, s, side, other_local)  which side of the face lc is on (0 = elem1) and, for a closed face, the other
//                 cell's tile-local index (-1: boundary face)
template <class Emits, class Kind, class PackInfo>
void compute_tile_order(int cell_count, int shift, bool by_cell, bool pack, Emits emits, Kind kind_of, PackInfo pinfo,
                        TileOrder &O) {
  struct Emit {
    int lc, s;
  };
  std::vector<Emit> group[3];
  std::vector<PackItem> pitems[3];
  for (int it = 0; it < 6 * cell_count; ++it) {
    const int lc = by_cell ? it / 6 : it % cell_count;
    const int s = by_cell ? it % 6 : it / cell_count;
    if (!emits(lc, s)) continue;
    const int kind = kind_of(lc, s);
    const bool cut = kind != 0;
    group[kind].push_back({lc, s});
    if (pack && kind < 2) {
      int side = 0, other_local = -1;
      pinfo(lc, s, side, other_local);
      const int own = shift + lc;
      PackItem pi;
      if (cut) {
        pi = {own, -1, side == 0 ? 2 : 3, 0};
      } else if (other_local < 0) {
        pi = {own, -1, 2, 0};
      } else {
        const int op = shift + other_local;
        pi = side == 0 ? PackItem{own, op, 0, 1} : PackItem{op, own, 0, 1};
      }
      pitems[cut].push_back(pi);
    }
  }
  std::vector<int> porder;
  for (int g = 0; g < 3; ++g) {
    if (pack && g < 2) {
      // the closed group follows the cut group in the kernel's work-item numbering
      pack_conflict_free(pitems[g], g == 0 ? (int)(group[1].size() & 15) : 0, porder);
    } else {
      porder.resize(group[g].size());
      for (size_t i = 0; i < group[g].size(); ++i) porder[i] = (int)i;
    }
    O.lc[g].resize(group[g].size());
    O.slot[g].resize(group[g].size());
    for (size_t q = 0; q < group[g].size(); ++q) {
      O.lc[g][q] = (uint16_t)group[g][porder[q]].lc;
      O.slot[g][q] = (uint8_t)group[g][porder[q]].s;
    }
  }
  if (!O.lc[2].empty() && ((O.lc[0].size() + O.lc[1].size()) & 1)) {
    const int dg = O.lc[0].empty() ? 1 : 0;
    O.lc[dg].push_back(O.lc[dg][0]);
    O.slot[dg].push_back(O.slot[dg][0]);
    O.dummy_group = dg;
  }
}

template <class Mesh>
int build_layout_impl(Mesh &mesh, const int tile_dims_in[3], bool with_tangents, bool defer_geometry, bool share_in,
                      HostLayout &L) {
  const int n_owned = mesh.n_owned, n_ghost = mesh.n_ghost;
  const long n_cells = mesh.n_cells;
  L = HostLayout();
  L.n_owned = n_owned;
  L.n_ghost = n_ghost;
  L.geom_components = with_tangents ? 12 : 6;
  L.geometry_deferred = defer_geometry;
  const bool share = share_in && !with_tangents;  // the STRICT kernels evaluate every tile face themselves
  L.share_cut_faces = share;
  L.stride = round_up(n_cells, 32);
  for (int d = 0; d < 3; ++d) L.tile_dims[d] = tile_dims_in[d] > 0 ? tile_dims_in[d] : 8;
  L.max_tile_cells = L.tile_dims[0] * L.tile_dims[1] * L.tile_dims[2];
  if (L.max_tile_cells > 4096) return ma_set_error(MA_ERR_INVALID, "tile_dims: at most 4096 cells per tile");

  const bool timing = getenv("MINIAERO_LAYOUT_TIMING") != nullptr;
  auto tprev = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "layout: %-28s %.3f s\n", what, std::chrono::duration<double>(now - tprev).count());
    tprev = now;
  };
  // ---- 2. spatial binning of owned cells
  mesh.prepare_bins();
  long nbin[3], ntile_d[3];
  for (int d = 0; d < 3; ++d) nbin[d] = mesh.nbins[d];
  for (int d = 0; d < 3; ++d) ntile_d[d] = (nbin[d] + L.tile_dims[d] - 1) / L.tile_dims[d];

  const char *order_env = getenv("MINIAERO_TILE_ORDER");  // experiment knob: "linear" = x-major tile order
  const bool linear_order = order_env && !strcmp(order_env, "linear");
  const char *sw_env = getenv("MINIAERO_CELL_SWIZZLE");  // experiment knob: 0 = plain z-fastest order inside a tile
  const bool swizzle = !with_tangents && L.tile_dims[0] == 4 && L.tile_dims[1] == 4 && L.tile_dims[2] == 8 &&
                       !(sw_env && sw_env[0] == '0');
  struct Key {
    uint64_t tile;
    uint32_t local;
    int cell;
  };
  auto tile_key_of = [&](long t0, long t1, long t2) -> uint64_t {
    return linear_order ? ((uint64_t)t0 * ntile_d[1] + t1) * ntile_d[2] + t2 : morton3(t0, t1, t2);
  };
  auto local_key_of = [&](long l0, long l1, long l2) -> uint32_t {
    uint32_t local = (uint32_t)((l0 * L.tile_dims[1] + l1) * L.tile_dims[2] + l2);
    if (swizzle) {
      // 4 x 4 x 8 bricks, z fastest: the cells of a brick SURFACE (the own cells of the cut faces) would share few
      // shared-memory banks (z surface: 2 of the 16 double-word residues, y surface: 8).  XOR-ing the low four bits
      // with a function of the 16-cell group index h = (x, y / 2) spreads every surface over all 16 residues while
      // aligned 8-cell z runs stay contiguous (the cut-face gathers keep their sector efficiency).
      const uint32_t h = local >> 4;
      local ^= ((h >> 1) & 3u) | ((h & 1u) << 2) | (((h >> 2) & 1u) << 3);
    }
    return local;
  };
  auto key_less = [](const Key &a, const Key &b) {
    if (a.tile != b.tile) return a.tile < b.tile;
    if (a.local != b.local) return a.local < b.local;
    return a.cell < b.cell;
  };
  std::vector<Key> keys((size_t)n_owned);
  if (Mesh::kStructured && (long)nbin[0] * nbin[1] * nbin[2] == (long)n_owned) {
    // a structured block: the sorted key array is written tile by tile (tiles sorted by key, the cells of a brick by
    // their local key), with no sort over the cells; cell (i, j, k) has id (i * ny + j) * nz + k
    struct TileRef {
      uint64_t key;
      int t[3];
      long first;
    };
    std::vector<TileRef> refs;
    refs.reserve((size_t)(ntile_d[0] * ntile_d[1] * ntile_d[2]));
    for (long a = 0; a < ntile_d[0]; ++a)
      for (long b = 0; b < ntile_d[1]; ++b)
        for (long c = 0; c < ntile_d[2]; ++c) refs.push_back({tile_key_of(a, b, c), {(int)a, (int)b, (int)c}, 0});
    std::sort(refs.begin(), refs.end(), [](const TileRef &x, const TileRef &y) { return x.key < y.key; });
    long next = 0;
    for (TileRef &r : refs) {
      r.first = next;
      long cnt = 1;
      for (int d = 0; d < 3; ++d) cnt *= std::min<long>(L.tile_dims[d], nbin[d] - (long)r.t[d] * L.tile_dims[d]);
      next += cnt;
    }
#pragma omp parallel for schedule(dynamic, 64)
    for (long ti = 0; ti < (long)refs.size(); ++ti) {
      const TileRef &r = refs[ti];
      long o[3], e[3];
      for (int d = 0; d < 3; ++d) {
        o[d] = (long)r.t[d] * L.tile_dims[d];
        e[d] = std::min<long>(L.tile_dims[d], nbin[d] - o[d]);
      }
      Key *out = keys.data() + r.first;
      long w = 0;
      for (long a = 0; a < e[0]; ++a)
        for (long b = 0; b < e[1]; ++b)
          for (long c = 0; c < e[2]; ++c)
            out[w++] = {r.key, local_key_of(a, b, c), (int)(((o[0] + a) * nbin[1] + (o[1] + b)) * nbin[2] + (o[2] + c))};
      std::sort(out, out + w, key_less);
    }
  } else {
#pragma omp parallel for schedule(static)
    for (long c = 0; c < n_owned; ++c) {
      long q[3], t[3], l[3];
      mesh.bin(c, q);
      for (int d = 0; d < 3; ++d) {
        t[d] = q[d] / L.tile_dims[d];
        l[d] = q[d] % L.tile_dims[d];
      }
      keys[c].tile = tile_key_of(t[0], t[1], t[2]);
      keys[c].local = local_key_of(l[0], l[1], l[2]);
      keys[c].cell = (int)c;
    }
#ifdef _OPENMP
    __gnu_parallel::sort(keys.begin(), keys.end(), key_less);
#else
    std::sort(keys.begin(), keys.end(), key_less);
#endif
  }

  lap("bin + sort cells");
  // ---- 3. cut tiles, classify (touches a ghost?), order interior tiles first, renumber
  struct RawTile {
    long first;
    int count;
    int boundary;
    int colour;  // parity of the tile's position in the tile lattice (shared cut faces: the flux pass it runs in)
  };
  std::vector<RawTile> raw;
  raw.reserve((size_t)n_owned / std::max(1, L.max_tile_cells / 2) + 16);
  for (long i = 0; i < n_owned;) {
    long j = i + 1;
    while (j < n_owned && keys[j].tile == keys[i].tile && (j - i) < L.max_tile_cells) ++j;
    raw.push_back({i, (int)(j - i), 0, 0});
    i = j;
  }
  const long n_tiles = (long)raw.size();
  if (share) {
#pragma omp parallel for schedule(static)
    for (long t = 0; t < n_tiles; ++t) {
      long q[3];
      mesh.bin(keys[raw[t].first].cell, q);
      raw[t].colour = (int)((q[0] / L.tile_dims[0] + q[1] / L.tile_dims[1] + q[2] / L.tile_dims[2]) & 1);
    }
  }
  if (n_ghost > 0) {
#pragma omp parallel for schedule(dynamic, 64)
    for (long t = 0; t < n_tiles; ++t) {
      int bnd = 0;
      for (long i = raw[t].first; i < raw[t].first + raw[t].count && !bnd; ++i) {
        const int c = keys[i].cell;
        for (int s = 0; s < 6; ++s)
          if (mesh.info(c, s).other >= n_owned) {
            bnd = 1;
            break;
          }
      }
      raw[t].boundary = bnd;
    }
  }
  std::vector<long> order;
  order.reserve(n_tiles);
  for (int cls = 0; cls < 4; ++cls) {  // launch classes: interior / boundary x first / second pass
    const size_t before = order.size();
    for (long t = 0; t < n_tiles; ++t)
      if ((raw[t].boundary ? 2 : 0) + raw[t].colour == cls) order.push_back(t);
    L.launch_count[cls] = (int)(order.size() - before);
  }
  L.n_interior_tiles = L.launch_count[0] + L.launch_count[1];
  L.n_tiles = (int)n_tiles;
  // launch class of every tile (in the new order) and, for shared cut faces, the tile of every owned cell
  std::vector<uint8_t> tile_launch((size_t)n_tiles, 0);
  std::vector<int> cell_tile;
  if (share) cell_tile.assign((size_t)n_owned, 0);
  for (long t = 0; t < n_tiles; ++t) L.max_tile_cells_real = std::max(L.max_tile_cells_real, raw[t].count);
  L.tiles.resize(n_tiles);
  L.new2old.resize(n_cells);
  L.old2new.resize(n_cells);
  {
    long next = 0;
    for (long k = 0; k < n_tiles; ++k) {
      L.tiles[k].cell_start = (int)next;
      L.tiles[k].cell_count = raw[order[k]].count;
      next += raw[order[k]].count;
    }
#pragma omp parallel for schedule(dynamic, 64)
    for (long k = 0; k < n_tiles; ++k) {
      const RawTile &r = raw[order[k]];
      tile_launch[k] = (uint8_t)((r.boundary ? 2 : 0) + r.colour);
      for (int i = 0; i < r.count; ++i) {
        const int oldc = keys[r.first + i].cell;
        const int newc = L.tiles[k].cell_start + i;
        L.new2old[newc] = oldc;
        L.old2new[oldc] = newc;
        if (share) cell_tile[newc] = (int)k;
      }
    }
    for (long g = n_owned; g < n_cells; ++g) L.new2old[g] = (int)g, L.old2new[g] = (int)g;
  }
  std::vector<Key>().swap(keys);

  lap("tiles + renumbering");
  // ---- 4. cell SoA
  if (!defer_geometry) {
    big_assign(L.cell_xyz, (size_t)3 * L.stride, 0.0);
    big_assign(L.cell_vol, (size_t)L.stride, 1.0);
  }
#pragma omp parallel for schedule(static)
  for (long c = 0; c < (defer_geometry ? 0 : n_cells); ++c) {
    const long o = L.new2old[c];
    double xyz[3], vol;
    mesh.cell_geometry(o, xyz, &vol);
    for (int d = 0; d < 3; ++d) L.cell_xyz[(size_t)d * L.stride + c] = xyz[d];
    L.cell_vol[c] = vol;
  }

  lap("cell geometry");
  // ---- 5. tile face lists.  A face is emitted by its in-tile cell with the larger tile-local index
  // (or by its only in-tile cell), in (cell, slot) order: the flux sweep then walks cells in order
  // and every cell's data is touched within a short window.
  L.slot_stride = round_up(n_owned, 32);
  big_assign(L.slot_face, (size_t)6 * L.slot_stride, (uint16_t)0);
  auto emits = [&](const TileInfo &T, int newc, int s) -> bool {
    const int oldc = L.new2old[newc];
    const int oth = mesh.info(oldc, s).other;
    if (oth < 0 || oth >= n_owned) return true;
    const int on = L.old2new[oth];
    if (on < T.cell_start || on >= T.cell_start + T.cell_count) return true;
    return on < newc;
  };
  // the face (new cell c, slot s) of tile T: 0 = the cell on the other side is in the tile too (or there is none),
  // 1 = cut face evaluated by this tile, 2 = cut face whose flux this tile imports (shared cut faces: the tile on the
  // other side runs in an earlier flux launch, evaluates the face and publishes the flux)
  auto cut_kind = [&](const TileInfo &T, int newc, int s) -> int {
    const int oldc = L.new2old[newc];
    const int oth = mesh.info(oldc, s).other;
    if (oth < 0) return 0;
    if (oth >= n_owned) return 1;
    const int on = L.old2new[oth];
    if (on >= T.cell_start && on < T.cell_start + T.cell_count) return 0;
    if (!share) return 1;
    return tile_launch[cell_tile[on]] < tile_launch[&T - L.tiles.data()] ? 2 : 1;
  };
  const char *fo_env = getenv("MINIAERO_FACE_ORDER");
  const bool by_cell = with_tangents || (fo_env && !strcmp(fo_env, "cell"));
  const bool pack = !by_cell && !(fo_env && !strcmp(fo_env, "slot"));
  // The tile-local face order: closed / boundary faces first, cut faces last (compute_tile_order).
  auto compute_order = [&](const TileInfo &T, TileOrder &O) {
    compute_tile_order(
        T.cell_count, T.cell_start & 1, by_cell, pack, [&](int lc, int sl) { return emits(T, T.cell_start + lc, sl); },
        [&](int lc, int sl) { return cut_kind(T, T.cell_start + lc, sl); },
        [&](int lc, int sl, int &side, int &other_local) {
          const SlotInfo si = mesh.info(L.new2old[T.cell_start + lc], sl);
          side = si.side;
          const int on = (si.other >= 0 && si.other < n_owned) ? L.old2new[si.other] : -1;
          other_local = (on >= T.cell_start && on < T.cell_start + T.cell_count) ? on - T.cell_start : -1;
        },
        O);
  };
  // structured blocks: tiles with the same key (StructuredAccess::tile_key) share one order
  std::unordered_map<uint64_t, std::shared_ptr<const TileOrder>> order_cache;
  std::unordered_map<std::string, std::shared_ptr<const TileOrder>> sig_cache;
  std::mutex order_mutex;
  auto order_of = [&](const TileInfo &T) -> std::shared_ptr<const TileOrder> {
    uint64_t key = 0;
    if (Mesh::kStructured && L.max_tile_cells < 4096)
      key = mesh.tile_key(L.new2old[T.cell_start], L.tile_dims, T.cell_count, T.cell_start & 1);
    if (key && share) {
      // which sides are evaluated here, which imported: part of the pattern (bit 2s: some cut face through slot
      // direction s is evaluated, bit 2s+1: some is imported; a side that does both is no brick side: no sharing of
      // the order)
      unsigned rel = 0;
      for (int lc = 0; lc < T.cell_count; ++lc)
        for (int sl = 0; sl < 6; ++sl) {
          const int kind = cut_kind(T, T.cell_start + lc, sl);
          if (kind) rel |= 1u << (2 * sl + (kind - 1));
        }
      for (int sl = 0; sl < 6; ++sl)
        if (((rel >> (2 * sl)) & 3u) == 3u) key = 0;
      if (key) key = key << 12 | rel;
    }
    if (key) {
      std::lock_guard<std::mutex> lock(order_mutex);
      auto it = order_cache.find(key);
      if (it != order_cache.end()) return it->second;
    }
    // No pattern key (a mesh handed over as arrays, an irregular tile): the order is a function of what the three
    // callbacks of compute_tile_order answer — tiles that answer alike (nearly all of a mesh with any regularity)
    // share one order.  The key is the answers themselves, so equal keys mean equal orders exactly.
    std::string sig;
    if (!key && T.cell_count <= 4096) {
      sig.resize(8 + (size_t)24 * T.cell_count);
      uint32_t *w = reinterpret_cast<uint32_t *>(&sig[0]);
      w[0] = (uint32_t)T.cell_count, w[1] = (uint32_t)(T.cell_start & 1);
      for (int lc = 0; lc < T.cell_count; ++lc)
        for (int sl = 0; sl < 6; ++sl) {
          const int c = T.cell_start + lc;
          uint32_t v = 0;
          if (emits(T, c, sl)) {
            const int kind = cut_kind(T, c, sl);
            const SlotInfo si = mesh.info(L.new2old[c], sl);
            const int on = (si.other >= 0 && si.other < n_owned) ? L.old2new[si.other] : -1;
            const int other_local = (on >= T.cell_start && on < T.cell_start + T.cell_count) ? on - T.cell_start : -1;
            v = 1u | (uint32_t)kind << 1 | (uint32_t)si.side << 3 | (uint32_t)(other_local + 1) << 4;
          }
          w[2 + 6 * lc + sl] = v;
        }
      std::lock_guard<std::mutex> lock(order_mutex);
      auto it = sig_cache.find(sig);
      if (it != sig_cache.end()) return it->second;
    }
    auto O = std::make_shared<TileOrder>();
    compute_order(T, *O);
    if (key || !sig.empty()) {
      std::lock_guard<std::mutex> lock(order_mutex);
      if (key)
        order_cache.emplace(key, O);
      else
        sig_cache.emplace(std::move(sig), O);
    }
    return O;
  };
  std::vector<long> fstart(n_tiles + 1, 0), hstart(n_tiles + 1, 0);
  int max_faces = 0, max_local = 0, max_halo = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(max : max_faces, max_local, max_halo)
  for (long k = 0; k < n_tiles; ++k) {
    const TileInfo &T = L.tiles[k];
    const std::shared_ptr<const TileOrder> O = order_of(T);
    const int cut = (int)(O->lc[1].size() + O->lc[2].size()), cnt = cut + (int)O->lc[0].size();
    L.tiles[k].face_count = cnt;
    L.tiles[k].cut_start = cnt - cut;
    L.tiles[k].n_eval = cnt - (int)O->lc[2].size();
    L.tiles[k].imp_area = O->lc[2].empty() ? -1 : 0;  // numbered below
    max_faces = std::max(max_faces, cnt);
    max_local = std::max(max_local, round_up((T.cell_start & 1) + T.cell_count, 2) + cut);
    max_halo = std::max(max_halo, cut);
  }
  lap("tile face orders");
  if (max_faces >= 16384) return ma_set_error(MA_ERR_INVALID, "tile has more than 16383 faces; use smaller tile_dims");
  if (max_local >= 0xFFF0 - 2) return ma_set_error(MA_ERR_INVALID, "tile has too many cells + cut faces; use smaller tile_dims");
  L.max_tile_faces = max_faces;
  L.max_tile_local = max_local;
  L.max_tile_halo = max_halo;
  // every tile owns halo_stride entries of tile_halo (unused ones are -1): the list of a tile sits at an address
  // computable from the tile index alone, so a CTA can fetch it together with — not after — its tile descriptor
  L.halo_stride = std::max(4, round_up(max_halo, 4));
  if ((long)n_tiles * L.halo_stride >= (1L << 31)) return ma_set_error(MA_ERR_INVALID, "more than 2^31 tile halo entries");
  long real = 0;
  for (long k = 0; k < n_tiles; ++k) {
    L.tiles[k].face_start = (int)fstart[k];
    L.tiles[k].halo_start = (int)(k * L.halo_stride);
    fstart[k + 1] = fstart[k] + round_up(L.tiles[k].face_count, 16);
    hstart[k + 1] = hstart[k] + (L.tiles[k].face_count - L.tiles[k].cut_start);
    real += L.tiles[k].face_count;
    if (fstart[k + 1] >= (1L << 31)) return ma_set_error(MA_ERR_INVALID, "more than 2^31 tile faces");
  }
  L.n_tile_faces = fstart[n_tiles];
  L.n_tile_faces_real = real;
  {
    int areas = 0, cap = 0;
    for (long k = 0; k < n_tiles; ++k)
      if (L.tiles[k].imp_area >= 0) {
        L.tiles[k].imp_area = areas++;
        cap = std::max(cap, L.tiles[k].face_count - L.tiles[k].n_eval);
      }
    L.n_import_areas = areas;
    L.import_capacity = round_up(cap, 2);
    if ((long)areas * 5 * L.import_capacity >= (1L << 31)) return ma_set_error(MA_ERR_INVALID, "more than 2^31 shared cut-face flux entries");
    if (share) big_assign(L.tile_pub, (size_t)n_tiles * L.halo_stride, -1);
  }
  const size_t NF = (size_t)L.n_tile_faces;
  if (defer_geometry)
    big_assign(L.face_code, NF, (uint32_t)0);
  else
    big_assign(L.face_geom, (size_t)L.geom_components * NF, 0.0);
  const int GX = with_tangents ? 9 : 3;  // first centroid component
  double frame_err = 0.0;
  big_assign(L.face_left, NF, 0);
  big_assign(L.face_right, NF, 0);
  big_assign(L.face_lr, NF, (uint32_t)0);
  big_assign(L.slot_nbr, (size_t)6 * L.slot_stride, (uint16_t)0xFFFF);
  big_assign(L.tile_halo, (size_t)n_tiles * L.halo_stride, -1);
  lap("face arrays allocated");
#pragma omp parallel for schedule(dynamic, 64) reduction(max : frame_err)
  for (long k = 0; k < n_tiles; ++k) {
    const TileInfo &T = L.tiles[k];
    const size_t fcp = (size_t)round_up(T.face_count, 16);
    const int shift = T.cell_start & 1, halo_base = round_up(shift + T.cell_count, 2);
    const std::shared_ptr<const TileOrder> O = order_of(T);
    for (int g = 0; g < 3; ++g)
    for (size_t q = 0; q < O->lc[g].size(); ++q) {
      const int c = T.cell_start + O->lc[g][q], s = O->slot[g][q];
      const int oldc = L.new2old[c];
      {
        const bool cut = g >= 1;
        const bool dummy = g == O->dummy_group && q + 1 == O->lc[g].size();  // repeats an evaluated face: no slot owns it
        const int e = (g == 0 ? 0 : g == 1 ? T.cut_start : T.n_eval) + (int)q;
        const SlotInfo si = mesh.info(oldc, s);
        const int side = si.side;
        const size_t j = (size_t)T.face_start + e;
        if (defer_geometry) {
          L.face_code[j] = mesh.face_code(oldc, s);
        } else {
          double fn[3], ft[3], fb[3], fx[3];
          mesh.face_geometry(oldc, s, fn, ft, fb, fx);
          for (int d = 0; d < 3; ++d) {
            if (with_tangents) {
              L.face_geom[(0 + d) * NF + j] = fn[d];
              L.face_geom[(3 + d) * NF + j] = ft[d];
              L.face_geom[(6 + d) * NF + j] = fb[d];
              L.face_geom[(GX + d) * NF + j] = fx[d];
            } else {
              const size_t base = (size_t)6 * T.face_start + e;
              L.face_geom[base + (0 + d) * fcp] = fn[d];
              L.face_geom[base + (3 + d) * fcp] = fx[d];
            }
          }
          if (!Mesh::kStructured) {  // how far (n/|n|, t, b/|n|) is from orthonormal (FAST arithmetic relies on it,
                                     // Face.C:81-96; the in-code mesh builds it that way)
            const double a2 = fn[0] * fn[0] + fn[1] * fn[1] + fn[2] * fn[2];
            const double an = std::sqrt(a2);
            const double tt = ft[0] * ft[0] + ft[1] * ft[1] + ft[2] * ft[2];
            const double bb = (fb[0] * fb[0] + fb[1] * fb[1] + fb[2] * fb[2]) / a2;
            const double nt = (fn[0] * ft[0] + fn[1] * ft[1] + fn[2] * ft[2]) / an;
            const double nb = (fn[0] * fb[0] + fn[1] * fb[1] + fn[2] * fb[2]) / a2;
            const double tb = (ft[0] * fb[0] + ft[1] * fb[1] + ft[2] * fb[2]) / an;
            double err = std::max(std::fabs(tt - 1.0), std::fabs(bb - 1.0));
            err = std::max(err, std::max(std::fabs(nt), std::max(std::fabs(nb), std::fabs(tb))));
            if (!(err <= frame_err)) frame_err = (err == err) ? err : 1e300;
          }
        }
        const int lc = c - T.cell_start;  // tile-local index of the emitting cell
        if (si.bc_type >= 0) {
          if (!dummy) L.slot_face[(size_t)s * L.slot_stride + c] = (uint16_t)(e | (1 << 14) | (side << 15));
          L.face_left[j] = c;
          L.face_right[j] = bc_code(si.bc_type);
          L.face_lr[j] = (uint32_t)(shift + lc) | ((uint32_t)(0xFFFF - si.bc_type) << 16);
        } else {
          if (!dummy) L.slot_face[(size_t)s * L.slot_stride + c] = (uint16_t)(e | (side << 15));
          const int oth_old = si.other;
          const int oth_new = L.old2new[oth_old];
          L.face_left[j] = side == 0 ? c : oth_new;
          L.face_right[j] = side == 0 ? oth_new : c;
          int oth_local;
          if (cut) {
            oth_local = halo_base + (e - T.cut_start);
            L.tile_halo[(size_t)T.halo_start + (e - T.cut_start)] = oth_new;
          } else {
            oth_local = shift + (oth_new - T.cell_start);
            const int os = si.other_slot;
            if (!dummy) {
              L.slot_face[(size_t)os * L.slot_stride + oth_new] = (uint16_t)(e | ((1 - side) << 15));
              L.slot_nbr[(size_t)os * L.slot_stride + oth_new] = (uint16_t)(shift + lc);
            }
          }
          if (!dummy) L.slot_nbr[(size_t)s * L.slot_stride + c] = (uint16_t)oth_local;
          L.face_lr[j] = side == 0 ? ((uint32_t)(shift + lc) | ((uint32_t)oth_local << 16))
                                   : ((uint32_t)oth_local | ((uint32_t)(shift + lc) << 16));
        }
      }
    }
  }

  L.max_frame_error = frame_err;
  lap("tile face lists");

  if (share) {
    // where every evaluated cut face publishes its flux: the import slot of the same face in the tile on the other
    // side, when that tile imports it (slot_face of the other cell is complete now)
    long bad_pub = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : bad_pub)
    for (long k = 0; k < n_tiles; ++k) {
      const TileInfo &T = L.tiles[k];
      const std::shared_ptr<const TileOrder> O = order_of(T);
      for (size_t q = 0; q < O->lc[1].size(); ++q) {
        if (O->dummy_group == 1 && q + 1 == O->lc[1].size()) continue;
        const int c = T.cell_start + O->lc[1][q], s = O->slot[1][q];
        const SlotInfo si = mesh.info(L.new2old[c], s);
        if (si.other < 0 || si.other >= n_owned) continue;
        const int on = L.old2new[si.other];
        const long k2 = cell_tile[on];
        if (!(tile_launch[k] < tile_launch[k2])) continue;  // the other tile evaluates the face itself
        const TileInfo &T2 = L.tiles[k2];
        const int e2 = L.slot_face[(size_t)si.other_slot * L.slot_stride + on] & 0x3fff;
        const int j2 = e2 - T2.n_eval;
        if (j2 < 0 || e2 >= T2.face_count || T2.imp_area < 0) {
          ++bad_pub;
          continue;
        }
        L.tile_pub[(size_t)T.halo_start + q] = T2.imp_area * 5 * L.import_capacity + j2;
      }
    }
    if (bad_pub) return ma_set_error(MA_ERR_INVALID, "shared cut faces: a published face is not in the importing tile's list (internal error)");
  }

  lap("publish lists");
  // ---- 6. halo lists (renumbered), grouped by peer
  if (n_ghost > 0) {
    long so = 0, ro = 0;
    for (int p = 0; p < mesh.num_ranks(); ++p) {
      const int sc = mesh.send_count(p), rc = mesh.recv_count(p);
      if (sc < 0 || rc < 0) return ma_set_error(MA_ERR_INVALID, "mesh: negative send/recv count");
      if (p == mesh.my_rank() || (sc == 0 && rc == 0)) {
        so += sc, ro += rc;
        continue;
      }
      L.peer_rank.push_back(p);
      L.peer_send_count.push_back(sc);
      L.peer_recv_count.push_back(rc);
      for (int i = 0; i < sc; ++i) {
        const int id = mesh.send_id(so + i);
        if (id < 0 || id >= n_owned) return ma_set_error(MA_ERR_INVALID, "mesh: send_local_ids out of range");
        L.send_ids.push_back(L.old2new[id]);
      }
      for (int i = 0; i < rc; ++i) {
        const int id = mesh.recv_id(ro + i);
        if (id < n_owned || id >= n_cells) return ma_set_error(MA_ERR_INVALID, "mesh: recv_local_ids must be ghosts");
        L.recv_ids.push_back(L.old2new[id]);
      }
      so += sc, ro += rc;
    }
  }
  return MA_OK;
}

}  // namespace

int build_layout(const ma_mesh &mesh, const int tile_dims[3], bool with_tangents, HostLayout &L, bool share_cut_faces) {
  ArrayAccess a(mesh);
  int rc = a.prepare();
  if (rc) return rc;
  return build_layout_impl(a, tile_dims, with_tangents, false, share_cut_faces, L);
}

int build_layout_structured(const ma_options &opt, int rank, int num_ranks, const int tile_dims[3], bool with_tangents,
                            bool defer_geometry, HostLayout &L, StructuredGrid *grid, bool share_cut_faces) {
  StructuredAccess a;
  int rc = a.prepare(opt, rank, num_ranks);
  if (rc) return rc;
  rc = build_layout_impl(a, tile_dims, with_tangents, defer_geometry, share_cut_faces, L);
  if (rc) return rc;
  if (grid) {
    grid->gen = a.g;
    grid->tables = std::move(a.tables);
    grid->gen.xs = grid->tables.xs.data(), grid->gen.ys = grid->tables.ys.data(), grid->gen.zs = grid->tables.zs.data();
  }
  return MA_OK;
}

// ---- topology plan for the device-side builder (layout.h: TopoPlan) -------------------------------------------------
// Mirrors build_layout_impl<StructuredAccess> step by step — same tile keys, same class order, same local cell order,
// the same compute_tile_order — but stops at O(tiles) + O(patterns); tests/test_gpu_topology.py compares every array
// the device then builds with the host builder's, bit for bit.
int build_topology_plan(const ma_options &opt, int rank, int num_ranks, const int tile_dims_in[3], bool share,
                        HostLayout &L, StructuredGrid *grid, TopoPlan &P) {
  StructuredAccess acc;
  int rc = acc.prepare(opt, rank, num_ranks);
  if (rc) return rc;
  const GridGen &g = acc.g;
  L = HostLayout();
  P = TopoPlan();
  const int n_owned = acc.n_owned, n_ghost = acc.n_ghost;
  const long n_cells = acc.n_cells;
  L.n_owned = n_owned, L.n_ghost = n_ghost;
  L.geom_components = 6;
  L.geometry_deferred = true;
  L.topology_on_device = true;
  L.share_cut_faces = share;
  L.stride = round_up(n_cells, 32);
  for (int d = 0; d < 3; ++d) L.tile_dims[d] = tile_dims_in[d] > 0 ? tile_dims_in[d] : 8;
  L.max_tile_cells = L.tile_dims[0] * L.tile_dims[1] * L.tile_dims[2];
  if (L.max_tile_cells > 4096) return ma_set_error(MA_ERR_INVALID, "tile_dims: at most 4096 cells per tile");
  if (L.tile_dims[0] > 255 || L.tile_dims[1] > 255 || L.tile_dims[2] > 255)
    return ma_set_error(MA_ERR_INVALID, "tile_dims: at most 255 cells per direction");
  for (int f = 0; f < 6; ++f) P.bc_of_face[f] = acc.bc_of_face[f];
  long nbin[3], ntile_d[3];
  for (int d = 0; d < 3; ++d) nbin[d] = g.b.n[d], ntile_d[d] = (nbin[d] + L.tile_dims[d] - 1) / L.tile_dims[d];
  const char *order_env = getenv("MINIAERO_TILE_ORDER");
  const bool linear_order = order_env && !strcmp(order_env, "linear");
  const char *sw_env = getenv("MINIAERO_CELL_SWIZZLE");
  const bool swizzle = L.tile_dims[0] == 4 && L.tile_dims[1] == 4 && L.tile_dims[2] == 8 && !(sw_env && sw_env[0] == '0');
  const char *fo_env = getenv("MINIAERO_FACE_ORDER");
  const bool by_cell = fo_env && !strcmp(fo_env, "cell");
  const bool pack = !by_cell && !(fo_env && !strcmp(fo_env, "slot"));
  auto local_key_of = [&](long l0, long l1, long l2) -> uint32_t {
    uint32_t local = (uint32_t)((l0 * L.tile_dims[1] + l1) * L.tile_dims[2] + l2);
    if (swizzle) {
      const uint32_t h = local >> 4;
      local ^= ((h >> 1) & 3u) | ((h & 1u) << 2) | (((h >> 2) & 1u) << 3);
    }
    return local;
  };
  // ---- tiles in key order, then by launch class
  struct Ref {
    uint64_t key;
    int t[3];
    int count, boundary, colour;
  };
  std::vector<Ref> refs;
  refs.reserve((size_t)(ntile_d[0] * ntile_d[1] * ntile_d[2]));
  for (long a = 0; a < ntile_d[0]; ++a)
    for (long b = 0; b < ntile_d[1]; ++b)
      for (long c = 0; c < ntile_d[2]; ++c) {
        Ref r;
        r.key = linear_order ? ((uint64_t)a * ntile_d[1] + b) * ntile_d[2] + c : morton3(a, b, c);
        r.t[0] = (int)a, r.t[1] = (int)b, r.t[2] = (int)c;
        r.count = 1, r.boundary = 0;
        for (int d = 0; d < 3; ++d) {
          const long o = (long)r.t[d] * L.tile_dims[d];
          const long e = std::min<long>(L.tile_dims[d], nbin[d] - o);
          r.count *= (int)e;
          if ((o == 0 && g.b.glo[d]) || (o + e == nbin[d] && g.b.ghi[d])) r.boundary = 1;
        }
        r.colour = share ? (int)((a + b + c) & 1) : 0;
        refs.push_back(r);
      }
  std::sort(refs.begin(), refs.end(), [](const Ref &x, const Ref &y) { return x.key < y.key; });
  const long n_tiles = (long)refs.size();
  std::vector<long> order;
  order.reserve(n_tiles);
  for (int cls = 0; cls < 4; ++cls) {
    const size_t before = order.size();
    for (long t = 0; t < n_tiles; ++t)
      if ((refs[t].boundary ? 2 : 0) + refs[t].colour == cls) order.push_back(t);
    L.launch_count[cls] = (int)(order.size() - before);
  }
  L.n_interior_tiles = L.launch_count[0] + L.launch_count[1];
  L.n_tiles = (int)n_tiles;
  L.tiles.resize(n_tiles);
  P.tile_pattern.assign(n_tiles, 0);
  P.tile_origin.assign((size_t)3 * n_tiles, 0);
  P.tile_nb.assign((size_t)6 * n_tiles, -1);
  P.tile_launch.assign(n_tiles, 0);
  std::vector<int> lattice_tile((size_t)(ntile_d[0] * ntile_d[1] * ntile_d[2]), -1);
  auto lat = [&](long a, long b, long c) { return (size_t)((a * ntile_d[1] + b) * ntile_d[2] + c); };
  {
    long next = 0;
    for (long k = 0; k < n_tiles; ++k) {
      const Ref &r = refs[order[k]];
      L.tiles[k].cell_start = (int)next;
      L.tiles[k].cell_count = r.count;
      next += r.count;
      L.max_tile_cells_real = std::max(L.max_tile_cells_real, r.count);
      for (int d = 0; d < 3; ++d) P.tile_origin[3 * k + d] = r.t[d] * L.tile_dims[d];
      P.tile_launch[k] = (uint8_t)((r.boundary ? 2 : 0) + r.colour);
      lattice_tile[lat(r.t[0], r.t[1], r.t[2])] = (int)k;
    }
  }
  for (long k = 0; k < n_tiles; ++k) {
    const Ref &r = refs[order[k]];
    for (int sl = 0; sl < 6; ++sl) {
      int dd[3];
      face_dir(sl, dd[0], dd[1], dd[2]);
      const int d = dd[0] ? 0 : dd[1] ? 1 : 2, dir = dd[d];
      const long nt = r.t[d] + dir;
      int nb;
      if (nt >= 0 && nt < ntile_d[d]) {
        long q[3] = {r.t[0], r.t[1], r.t[2]};
        q[d] = nt;
        nb = lattice_tile[lat(q[0], q[1], q[2])];
      } else {
        nb = (dir < 0 ? g.b.glo[d] : g.b.ghi[d]) ? -2 : -1;
      }
      P.tile_nb[6 * k + sl] = nb;
    }
  }
  // ---- patterns
  std::unordered_map<uint64_t, int> pattern_of_key;
  auto kind_of_side = [&](long k, int sl) -> int {  // of the cut faces through slot direction sl: 0 none, 1 evaluated, 2 imported
    const int nb = P.tile_nb[6 * k + sl];
    if (nb == -1) return 0;
    if (nb == -2) return 1;
    return (share && P.tile_launch[nb] < P.tile_launch[k]) ? 2 : 1;
  };
  for (long k = 0; k < n_tiles; ++k) {
    const TileInfo &T = L.tiles[k];
    const int *o = &P.tile_origin[3 * k];
    int ext[3];
    for (int d = 0; d < 3; ++d) ext[d] = (int)std::min<long>(L.tile_dims[d], nbin[d] - o[d]);
    uint64_t key = acc.tile_key((int)g.cell_id(o[0], o[1], o[2]), L.tile_dims, T.cell_count, T.cell_start & 1);
    unsigned rel = 0;
    for (int sl = 0; sl < 6; ++sl) {
      const int kd = kind_of_side(k, sl);
      if (kd) rel |= 1u << (2 * sl + (kd - 1));
    }
    key = key << 12 | rel;
    auto it = pattern_of_key.find(key);
    if (it != pattern_of_key.end()) {
      P.tile_pattern[k] = it->second;
      continue;
    }
    TopoPattern pat;
    for (int d = 0; d < 3; ++d) pat.ext[d] = ext[d];
    pat.cell_count = T.cell_count;
    struct LK {
      uint32_t local;
      int cell;
      uint32_t abc;
    };
    std::vector<LK> lk;
    lk.reserve(T.cell_count);
    for (int a = 0; a < ext[0]; ++a)
      for (int b = 0; b < ext[1]; ++b)
        for (int c = 0; c < ext[2]; ++c)
          lk.push_back({local_key_of(a, b, c), (int)g.cell_id(o[0] + a, o[1] + b, o[2] + c),
                        (uint32_t)a | (uint32_t)b << 8 | (uint32_t)c << 16});
    std::sort(lk.begin(), lk.end(), [](const LK &x, const LK &y) { return x.local != y.local ? x.local < y.local : x.cell < y.cell; });
    pat.cell_abc.resize(T.cell_count);
    pat.rank_of.assign((size_t)ext[0] * ext[1] * ext[2], 0);
    for (int lc = 0; lc < T.cell_count; ++lc) {
      pat.cell_abc[lc] = lk[lc].abc;
      const int a = lk[lc].abc & 255, b = (lk[lc].abc >> 8) & 255, c = lk[lc].abc >> 16;
      pat.rank_of[((size_t)a * ext[1] + b) * ext[2] + c] = (uint16_t)lc;
    }
    // the neighbour of (lc, slot): tile-local index when it is in the brick, else -1
    auto inside = [&](int lc, int sl) -> int {
      int da, db, dc;
      face_dir(sl, da, db, dc);
      const int a = (int)(pat.cell_abc[lc] & 255) + da, b = (int)((pat.cell_abc[lc] >> 8) & 255) + db,
                c = (int)(pat.cell_abc[lc] >> 16) + dc;
      if (a < 0 || b < 0 || c < 0 || a >= ext[0] || b >= ext[1] || c >= ext[2]) return -1;
      return pat.rank_of[((size_t)a * ext[1] + b) * ext[2] + c];
    };
    TileOrder O;
    compute_tile_order(
        T.cell_count, T.cell_start & 1, by_cell, pack,
        [&](int lc, int sl) {
          const int ol = inside(lc, sl);
          return ol < 0 ? true : ol < lc;
        },
        [&](int lc, int sl) { return inside(lc, sl) >= 0 ? 0 : kind_of_side(k, sl); },
        [&](int lc, int sl, int &side, int &other_local) {
          const int a = (int)(pat.cell_abc[lc] & 255), b = (int)((pat.cell_abc[lc] >> 8) & 255), c = (int)(pat.cell_abc[lc] >> 16);
          side = acc.info((int)g.cell_id(o[0] + a, o[1] + b, o[2] + c), sl).side;
          other_local = inside(lc, sl);
        },
        O);
    pat.cut_start = (int)O.lc[0].size();
    pat.n_eval = pat.cut_start + (int)O.lc[1].size();
    pat.face_count = pat.n_eval + (int)O.lc[2].size();
    pat.dummy_face = O.dummy_group < 0 ? -1 : (O.dummy_group == 0 ? pat.cut_start - 1 : pat.n_eval - 1);
    for (int gi = 0; gi < 3; ++gi)
      for (size_t q = 0; q < O.lc[gi].size(); ++q) {
        pat.face_lc.push_back(O.lc[gi][q]);
        pat.face_slot.push_back(O.slot[gi][q]);
      }
    const int id = (int)P.patterns.size();
    P.patterns.push_back(std::move(pat));
    pattern_of_key.emplace(key, id);
    P.tile_pattern[k] = id;
  }
  // ---- descriptors, sizes
  int max_faces = 0, max_local = 0, max_halo = 0;
  for (long k = 0; k < n_tiles; ++k) {
    const TopoPattern &pat = P.patterns[P.tile_pattern[k]];
    TileInfo &T = L.tiles[k];
    T.face_count = pat.face_count, T.cut_start = pat.cut_start, T.n_eval = pat.n_eval;
    T.imp_area = pat.face_count > pat.n_eval ? 0 : -1;
    const int cut = pat.face_count - pat.cut_start;
    max_faces = std::max(max_faces, pat.face_count);
    max_local = std::max(max_local, round_up((T.cell_start & 1) + T.cell_count, 2) + cut);
    max_halo = std::max(max_halo, cut);
  }
  if (max_faces >= 16384) return ma_set_error(MA_ERR_INVALID, "tile has more than 16383 faces; use smaller tile_dims");
  if (max_local >= 0xFFF0 - 2) return ma_set_error(MA_ERR_INVALID, "tile has too many cells + cut faces; use smaller tile_dims");
  L.max_tile_faces = max_faces, L.max_tile_local = max_local, L.max_tile_halo = max_halo;
  L.halo_stride = std::max(4, round_up(max_halo, 4));
  if ((long)n_tiles * L.halo_stride >= (1L << 31)) return ma_set_error(MA_ERR_INVALID, "more than 2^31 tile halo entries");
  {
    long fstart = 0, real = 0;
    int areas = 0, cap = 0;
    for (long k = 0; k < n_tiles; ++k) {
      TileInfo &T = L.tiles[k];
      T.face_start = (int)fstart;
      T.halo_start = (int)(k * L.halo_stride);
      fstart += round_up(T.face_count, 16);
      real += T.face_count;
      if (fstart >= (1L << 31)) return ma_set_error(MA_ERR_INVALID, "more than 2^31 tile faces");
      if (T.imp_area >= 0) {
        T.imp_area = areas++;
        cap = std::max(cap, T.face_count - T.n_eval);
      }
    }
    L.n_tile_faces = fstart, L.n_tile_faces_real = real;
    L.n_import_areas = areas;
    L.import_capacity = round_up(cap, 2);
    if ((long)areas * 5 * L.import_capacity >= (1L << 31)) return ma_set_error(MA_ERR_INVALID, "more than 2^31 shared cut-face flux entries");
  }
  L.slot_stride = round_up(n_owned, 32);
  // ---- exchange lists in renumbered ids (the renumbering itself happens on the device)
  auto old2new = [&](int old) -> int {
    if (old >= n_owned) return old;
    int i, j, k2;
    g.cell_ijk(old, i, j, k2);
    const long t0 = i / L.tile_dims[0], t1 = j / L.tile_dims[1], t2 = k2 / L.tile_dims[2];
    const int k = lattice_tile[lat(t0, t1, t2)];
    const TopoPattern &pat = P.patterns[P.tile_pattern[k]];
    const int a = i - P.tile_origin[3 * k], b = j - P.tile_origin[3 * k + 1], c = k2 - P.tile_origin[3 * k + 2];
    return L.tiles[k].cell_start + pat.rank_of[((size_t)a * pat.ext[1] + b) * pat.ext[2] + c];
  };
  if (n_ghost > 0) {
    long so = 0, ro = 0;
    for (int p = 0; p < acc.num_ranks(); ++p) {
      const int sc = acc.send_count(p), rcnt = acc.recv_count(p);
      if (p == acc.my_rank() || (sc == 0 && rcnt == 0)) {
        so += sc, ro += rcnt;
        continue;
      }
      L.peer_rank.push_back(p);
      L.peer_send_count.push_back(sc);
      L.peer_recv_count.push_back(rcnt);
      for (int i = 0; i < sc; ++i) L.send_ids.push_back(old2new(acc.send_id(so + i)));
      for (int i = 0; i < rcnt; ++i) L.recv_ids.push_back(old2new(acc.recv_id(ro + i)));
      so += sc, ro += rcnt;
    }
  }
  if (grid) {
    grid->gen = acc.g;
    grid->tables = std::move(acc.tables);
    grid->gen.xs = grid->tables.xs.data(), grid->gen.ys = grid->tables.ys.data(), grid->gen.zs = grid->tables.zs.data();
  }
  return MA_OK;
}

}  // namespace ma
