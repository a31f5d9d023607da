// See layout.h.  Everything here is one-time host setup (the analogue of the reference's
// copy_faces / copy_cell_data, Faces.h:88-145, Cells.h:126-146); it is O(cells) and OpenMP-parallel.
#include "layout.h"

#include <algorithm>
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

}  // namespace

int build_layout(const ma_mesh &mesh, const int tile_dims_in[3], bool with_tangents, HostLayout &L) {
  const int n_owned = mesh.num_owned_cells, n_ghost = mesh.num_ghosts;
  const long n_cells = (long)n_owned + n_ghost;
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

  L = HostLayout();
  L.n_owned = n_owned;
  L.n_ghost = n_ghost;
  L.geom_components = with_tangents ? 12 : 6;
  L.stride = round_up(n_cells, 32);
  for (int d = 0; d < 3; ++d) L.tile_dims[d] = tile_dims_in[d] > 0 ? tile_dims_in[d] : 8;
  L.max_tile_cells = L.tile_dims[0] * L.tile_dims[1] * L.tile_dims[2];
  if (L.max_tile_cells > 4096) return ma_set_error(MA_ERR_INVALID, "tile_dims: at most 4096 cells per tile");

  // ---- 1. cell -> face table over owned cells: ref = global_face*2 + side, global face numbering is
  // internal faces first, then the boundary sets in order.
  const long n_int = mesh.internal_faces.nfaces;
  std::vector<long> set_base(mesh.num_boundary_sets + 1, n_int);
  for (int b = 0; b < mesh.num_boundary_sets; ++b) set_base[b + 1] = set_base[b] + mesh.boundary_faces[b].nfaces;
  const long n_faces_all = set_base[mesh.num_boundary_sets];
  if (n_faces_all >= (1L << 31)) return ma_set_error(MA_ERR_INVALID, "mesh: more than 2^31 faces");
  const uint32_t kNone = 0xFFFFFFFFu;
  std::vector<uint32_t> cf((size_t)n_owned * 6, kNone);
  int bad = 0;
  {
    const int *conn = mesh.internal_faces.face_cell_conn, *slot = mesh.internal_faces.cell_flux_index;
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (long f = 0; f < n_int; ++f) {
      for (int side = 0; side < 2; ++side) {
        const int c = conn[2 * f + side], s = slot[2 * f + side];
        if (c < 0 || c >= n_cells || s < 0 || s > 5) {
          ++bad;
          continue;
        }
        if (c < n_owned) cf[(size_t)c * 6 + s] = (uint32_t)(f * 2 + side);
      }
    }
    for (int b = 0; b < mesh.num_boundary_sets; ++b) {
      const ma_faces &F = mesh.boundary_faces[b];
      const long base = set_base[b];
#pragma omp parallel for schedule(static) reduction(+ : bad)
      for (long f = 0; f < F.nfaces; ++f) {
        const int c = F.face_cell_conn[2 * f], s = F.cell_flux_index[2 * f];
        if (c < 0 || c >= n_owned || s < 0 || s > 5) {
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
  auto face_src = [&](uint32_t ref) {
    const long g = ref >> 1;
    FaceSrc s;
    if (g < n_int) {
      s.f = &mesh.internal_faces, s.index = (int)g, s.bc_type = -1;
    } else {
      int b = 0;
      while (g >= set_base[b + 1]) ++b;
      s.f = &mesh.boundary_faces[b], s.index = (int)(g - set_base[b]), s.bc_type = mesh.boundary_type[b];
    }
    return s;
  };
  // the cell on the other side of (owned) cell c's face `ref`; -1 for a boundary face
  auto other_cell = [&](uint32_t ref) -> int {
    const long g = ref >> 1;
    if (g >= n_int) return -1;
    return mesh.internal_faces.face_cell_conn[2 * g + (1 - (int)(ref & 1))];
  };

  // ---- 2. spatial binning of owned cells.  The mean centroid spacing along each axis is taken over
  // internal faces whose cell-to-cell vector is dominated by that axis; for a structured block this
  // recovers (i,j,k) exactly, for a general hex mesh it only has to give compact tiles.
  const double *xc = mesh.cell_coordinates;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
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
#pragma omp parallel for schedule(static) reduction(+ : s0, s1, s2, c0, c1, c2)
    for (long f = 0; f < n_int; ++f) {
      const int l = conn[2 * f], r = conn[2 * f + 1];
      if (l >= n_owned || r >= n_owned) continue;
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
  double h[3];
  long nbin[3];
  for (int d = 0; d < 3; ++d) {
    h[d] = hcnt[d] ? hsum[d] / (double)hcnt[d] : 0.0;
    if (!(h[d] > 0.0) || !((hi[d] - lo[d]) / h[d] < 1e9)) h[d] = (hi[d] - lo[d]) + 1.0;  // one bin
    nbin[d] = (long)std::floor((hi[d] - lo[d]) / h[d] + 0.5) + 1;
  }
  long ntile_d[3];
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
  std::vector<Key> keys((size_t)n_owned);
#pragma omp parallel for schedule(static)
  for (long c = 0; c < n_owned; ++c) {
    long q[3], t[3], l[3];
    for (int d = 0; d < 3; ++d) {
      q[d] = (long)std::floor((xc[3 * c + d] - lo[d]) / h[d] + 0.5);
      q[d] = std::max(0L, std::min(q[d], nbin[d] - 1));
      t[d] = q[d] / L.tile_dims[d];
      l[d] = q[d] % L.tile_dims[d];
    }
    keys[c].tile = linear_order ? ((uint64_t)t[0] * ntile_d[1] + t[1]) * ntile_d[2] + t[2] : morton3(t[0], t[1], t[2]);
    uint32_t local = (uint32_t)((l[0] * L.tile_dims[1] + l[1]) * L.tile_dims[2] + l[2]);
    if (swizzle) {
      // 4 x 4 x 8 bricks, z fastest: the cells of a brick SURFACE (the own cells of the cut faces) would share few
      // shared-memory banks (z surface: 2 of the 16 double-word residues, y surface: 8).  XOR-ing the low four bits
      // with a function of the 16-cell group index h = (x, y / 2) spreads every surface over all 16 residues while
      // aligned 8-cell z runs stay contiguous (the cut-face gathers keep their sector efficiency).
      const uint32_t h = local >> 4;
      local ^= ((h >> 1) & 3u) | ((h & 1u) << 2) | (((h >> 2) & 1u) << 3);
    }
    keys[c].local = local;
    keys[c].cell = (int)c;
  }
  auto key_less = [](const Key &a, const Key &b) {
    if (a.tile != b.tile) return a.tile < b.tile;
    if (a.local != b.local) return a.local < b.local;
    return a.cell < b.cell;
  };
#ifdef _OPENMP
  __gnu_parallel::sort(keys.begin(), keys.end(), key_less);
#else
  std::sort(keys.begin(), keys.end(), key_less);
#endif

  // ---- 3. cut tiles, classify (touches a ghost?), order interior tiles first, renumber
  struct RawTile {
    long first;
    int count;
    int boundary;
  };
  std::vector<RawTile> raw;
  raw.reserve((size_t)n_owned / std::max(1, L.max_tile_cells / 2) + 16);
  for (long i = 0; i < n_owned;) {
    long j = i + 1;
    while (j < n_owned && keys[j].tile == keys[i].tile && (j - i) < L.max_tile_cells) ++j;
    raw.push_back({i, (int)(j - i), 0});
    i = j;
  }
  const long n_tiles = (long)raw.size();
  if (n_ghost > 0) {
#pragma omp parallel for schedule(dynamic, 64)
    for (long t = 0; t < n_tiles; ++t) {
      int bnd = 0;
      for (long i = raw[t].first; i < raw[t].first + raw[t].count && !bnd; ++i) {
        const int c = keys[i].cell;
        for (int s = 0; s < 6; ++s)
          if (other_cell(cf[(size_t)c * 6 + s]) >= n_owned) {
            bnd = 1;
            break;
          }
      }
      raw[t].boundary = bnd;
    }
  }
  std::vector<long> order;
  order.reserve(n_tiles);
  for (long t = 0; t < n_tiles; ++t)
    if (!raw[t].boundary) order.push_back(t);
  L.n_interior_tiles = (int)order.size();
  for (long t = 0; t < n_tiles; ++t)
    if (raw[t].boundary) order.push_back(t);
  L.n_tiles = (int)n_tiles;
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
      for (int i = 0; i < r.count; ++i) {
        const int oldc = keys[r.first + i].cell;
        const int newc = L.tiles[k].cell_start + i;
        L.new2old[newc] = oldc;
        L.old2new[oldc] = newc;
      }
    }
    for (long g = n_owned; g < n_cells; ++g) L.new2old[g] = (int)g, L.old2new[g] = (int)g;
  }
  std::vector<Key>().swap(keys);

  // ---- 4. cell SoA
  L.cell_xyz.assign((size_t)3 * L.stride, 0.0);
  L.cell_vol.assign((size_t)L.stride, 1.0);
#pragma omp parallel for schedule(static)
  for (long c = 0; c < n_cells; ++c) {
    const long o = L.new2old[c];
    for (int d = 0; d < 3; ++d) L.cell_xyz[(size_t)d * L.stride + c] = xc[3 * o + d];
    L.cell_vol[c] = mesh.cell_volumes[o];
  }

  // ---- 5. tile face lists.  A face is emitted by its in-tile cell with the larger tile-local index
  // (or by its only in-tile cell), in (cell, slot) order: the flux sweep then walks cells in order
  // and every cell's data is touched within a short window.
  L.slot_stride = round_up(n_owned, 32);
  L.slot_face.assign((size_t)6 * L.slot_stride, 0);
  auto emits = [&](const TileInfo &T, int newc, int s) -> bool {
    const int oldc = L.new2old[newc];
    const int oth = other_cell(cf[(size_t)oldc * 6 + s]);
    if (oth < 0 || oth >= n_owned) return true;
    const int on = L.old2new[oth];
    if (on < T.cell_start || on >= T.cell_start + T.cell_count) return true;
    return on < newc;
  };
  // is the cell on the other side of (new cell c, slot s) outside the tile?  (cut face)
  auto is_cut = [&](const TileInfo &T, int newc, int s) -> bool {
    const int oldc = L.new2old[newc];
    const int oth = other_cell(cf[(size_t)oldc * 6 + s]);
    if (oth < 0) return false;
    if (oth >= n_owned) return true;
    const int on = L.old2new[oth];
    return on < T.cell_start || on >= T.cell_start + T.cell_count;
  };
  std::vector<long> fstart(n_tiles + 1, 0), hstart(n_tiles + 1, 0);
  int max_faces = 0, max_local = 0, max_halo = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(max : max_faces, max_local, max_halo)
  for (long k = 0; k < n_tiles; ++k) {
    const TileInfo &T = L.tiles[k];
    int cnt = 0, cut = 0;
    for (int c = T.cell_start; c < T.cell_start + T.cell_count; ++c)
      for (int s = 0; s < 6; ++s) {
        if (!emits(T, c, s)) continue;
        ++cnt;
        cut += is_cut(T, c, s) ? 1 : 0;
      }
    L.tiles[k].face_count = cnt;
    L.tiles[k].cut_start = cnt - cut;
    max_faces = std::max(max_faces, cnt);
    max_local = std::max(max_local, round_up((T.cell_start & 1) + T.cell_count, 2) + cut);
    max_halo = std::max(max_halo, cut);
  }
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
  const size_t NF = (size_t)L.n_tile_faces;
  L.face_geom.assign((size_t)L.geom_components * NF, 0.0);
  const int GX = with_tangents ? 9 : 3;  // first centroid component
  double frame_err = 0.0;
  L.face_left.assign(NF, 0);
  L.face_right.assign(NF, 0);
  L.face_lr.assign(NF, 0);
  L.slot_nbr.assign((size_t)6 * L.slot_stride, 0xFFFF);
  L.tile_halo.assign((size_t)n_tiles * L.halo_stride, -1);
  const char *fo_env = getenv("MINIAERO_FACE_ORDER");
  const bool by_cell = with_tangents || (fo_env && !strcmp(fo_env, "cell"));
  const bool pack = !by_cell && !(fo_env && !strcmp(fo_env, "slot"));
#pragma omp parallel for schedule(dynamic, 64) reduction(max : frame_err)
  for (long k = 0; k < n_tiles; ++k) {
    const TileInfo &T = L.tiles[k];
    const size_t fcp = (size_t)round_up(T.face_count, 16);
    const int shift = T.cell_start & 1, halo_base = round_up(shift + T.cell_count, 2);
    // closed / boundary faces first, cut faces last.  Inside each group the faces are listed by slot (direction),
    // then by cell — consecutive faces read consecutive own cells and consecutive neighbours — and, for the staged
    // FAST kernels, that list is re-packed half-warp by half-warp so that the 16 lanes of a shared-memory wavefront
    // read 16 different banks (pack_conflict_free; MINIAERO_FACE_ORDER=slot keeps the plain list).
    // MINIAERO_FACE_ORDER=cell (and STRICT) keeps (cell, slot) order (the gather kernels' L1 locality).
    struct Emit {
      int c, s;
    };
    std::vector<Emit> group[2];  // 0 closed / boundary, 1 cut
    std::vector<PackItem> pitems[2];
    for (int it = 0; it < 6 * T.cell_count; ++it) {
      const int c = T.cell_start + (by_cell ? it / 6 : it % T.cell_count);
      const int s = by_cell ? it % 6 : it / T.cell_count;
      if (!emits(T, c, s)) continue;
      const bool cut = is_cut(T, c, s);
      group[cut].push_back({c, s});
      if (pack) {
        const uint32_t ref = cf[(size_t)L.new2old[c] * 6 + s];
        const int side = (int)(ref & 1), own = shift + (c - T.cell_start);
        const int oth = other_cell(ref);
        PackItem pi;
        if (cut) {
          pi = {own, -1, side == 0 ? 2 : 3, 0};
        } else if (oth < 0) {
          pi = {own, -1, 2, 0};
        } else {
          const int op = shift + (L.old2new[oth] - T.cell_start);
          pi = side == 0 ? PackItem{own, op, 0, 1} : PackItem{op, own, 0, 1};
        }
        pitems[cut].push_back(pi);
      }
    }
    std::vector<int> porder[2];
    for (int g = 0; g < 2; ++g) {
      if (pack) {
        // the closed group follows the cut group in the kernel's work-item numbering
        pack_conflict_free(pitems[g], g == 0 ? (int)(group[1].size() & 15) : 0, porder[g]);
      } else {
        porder[g].resize(group[g].size());
        for (size_t i = 0; i < group[g].size(); ++i) porder[g][i] = (int)i;
      }
    }
    for (int g = 0; g < 2; ++g)
    for (size_t q = 0; q < group[g].size(); ++q) {
      const int c = group[g][porder[g][q]].c, s = group[g][porder[g][q]].s;
      const int oldc = L.new2old[c];
      {
        const bool cut = g == 1;
        const int e = (cut ? T.cut_start : 0) + (int)q;
        const uint32_t ref = cf[(size_t)oldc * 6 + s];
        const int side = (int)(ref & 1);
        const FaceSrc src = face_src(ref);
        const size_t j = (size_t)T.face_start + e;
        const size_t fi = (size_t)src.index;
        const double *fn = src.f->face_normal + 3 * fi, *ft = src.f->face_tangent + 3 * fi,
                     *fb = src.f->face_binormal + 3 * fi;
        for (int d = 0; d < 3; ++d) {
          if (with_tangents) {
            L.face_geom[(0 + d) * NF + j] = fn[d];
            L.face_geom[(3 + d) * NF + j] = ft[d];
            L.face_geom[(6 + d) * NF + j] = fb[d];
            L.face_geom[(GX + d) * NF + j] = src.f->coordinates[3 * fi + d];
          } else {
            const size_t base = (size_t)6 * T.face_start + e;
            L.face_geom[base + (0 + d) * fcp] = fn[d];
            L.face_geom[base + (3 + d) * fcp] = src.f->coordinates[3 * fi + d];
          }
        }
        {  // how far (n/|n|, t, b/|n|) is from orthonormal (FAST arithmetic relies on it, Face.C:81-96)
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
        const int lc = c - T.cell_start;  // tile-local index of the emitting cell
        if (src.bc_type >= 0) {
          L.slot_face[(size_t)s * L.slot_stride + c] = (uint16_t)(e | (1 << 14) | (side << 15));
          L.face_left[j] = c;
          L.face_right[j] = bc_code(src.bc_type);
          L.face_lr[j] = (uint32_t)(shift + lc) | ((uint32_t)(0xFFFF - src.bc_type) << 16);
        } else {
          L.slot_face[(size_t)s * L.slot_stride + c] = (uint16_t)(e | (side << 15));
          const int oth_old = src.f->face_cell_conn[2 * fi + (1 - side)];
          const int oth_new = L.old2new[oth_old];
          L.face_left[j] = side == 0 ? c : oth_new;
          L.face_right[j] = side == 0 ? oth_new : c;
          int oth_local;
          if (cut) {
            oth_local = halo_base + (e - T.cut_start);
            L.tile_halo[(size_t)T.halo_start + (e - T.cut_start)] = oth_new;
          } else {
            oth_local = shift + (oth_new - T.cell_start);
            const int os = src.f->cell_flux_index[2 * fi + (1 - side)];
            L.slot_face[(size_t)os * L.slot_stride + oth_new] = (uint16_t)(e | ((1 - side) << 15));
            L.slot_nbr[(size_t)os * L.slot_stride + oth_new] = (uint16_t)(shift + lc);
          }
          L.slot_nbr[(size_t)s * L.slot_stride + c] = (uint16_t)oth_local;
          L.face_lr[j] = side == 0 ? ((uint32_t)(shift + lc) | ((uint32_t)oth_local << 16))
                                   : ((uint32_t)oth_local | ((uint32_t)(shift + lc) << 16));
        }
      }
    }
  }

  L.max_frame_error = frame_err;

  // ---- 6. halo lists (renumbered), grouped by peer
  if (n_ghost > 0) {
    long so = 0, ro = 0;
    for (int p = 0; p < mesh.num_ranks; ++p) {
      const int sc = mesh.send_count[p], rc = mesh.recv_count[p];
      if (sc < 0 || rc < 0) return ma_set_error(MA_ERR_INVALID, "mesh: negative send/recv count");
      if (p == mesh.my_rank || (sc == 0 && rc == 0)) {
        so += sc, ro += rc;
        continue;
      }
      L.peer_rank.push_back(p);
      L.peer_send_count.push_back(sc);
      L.peer_recv_count.push_back(rc);
      for (int i = 0; i < sc; ++i) {
        const int id = mesh.send_local_ids[so + i];
        if (id < 0 || id >= n_owned) return ma_set_error(MA_ERR_INVALID, "mesh: send_local_ids out of range");
        L.send_ids.push_back(L.old2new[id]);
      }
      for (int i = 0; i < rc; ++i) {
        const int id = mesh.recv_local_ids[ro + i];
        if (id < n_owned || id >= n_cells) return ma_set_error(MA_ERR_INVALID, "mesh: recv_local_ids must be ghosts");
        L.recv_ids.push_back(L.old2new[id]);
      }
      so += sc, ro += rc;
    }
  }
  return MA_OK;
}

}  // namespace ma
