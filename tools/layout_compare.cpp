// Host-side check (no GPU): build_layout_structured() — the layout of one block of the in-code mesh straight from
// (i, j, k) — against build_layout() of the reference-format arrays ma_mesh_generate() makes for the same block.
// Every array of the two HostLayouts must be identical, bit for bit.
//   layout_compare NX NY NZ PROBLEM_TYPE ANGLE RANK NRANKS [tx ty tz] [strict]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "layout.h"
#include "miniaero_b200.h"

template <class T, class A, class B>
static long diff(const char *name, const std::vector<T, A> &a, const std::vector<T, B> &b) {
  long bad = a.size() != b.size();
  if (!bad) bad = a.empty() ? 0 : memcmp(a.data(), b.data(), a.size() * sizeof(T)) != 0;
  if (bad) printf("  DIFFERENT: %s (%zu vs %zu entries)\n", name, a.size(), b.size());
  return bad;
}

int main(int argc, char **argv) {
  if (argc < 8) return printf("usage: layout_compare NX NY NZ PTYPE ANGLE RANK NRANKS [tx ty tz] [strict]\n"), 2;
  ma_options opt;
  ma_options_default(&opt);
  opt.nx = atoi(argv[1]), opt.ny = atoi(argv[2]), opt.nz = atoi(argv[3]);
  opt.problem_type = atoi(argv[4]);
  opt.angle = atof(argv[5]);
  const int rank = atoi(argv[6]), nranks = atoi(argv[7]);
  opt.lx = opt.problem_type == 0 ? 0.3048 : 2.0, opt.ly = opt.problem_type == 1 ? 0.002 : 1.3, opt.lz = 0.9;
  int td[3] = {argc > 10 ? atoi(argv[8]) : 4, argc > 10 ? atoi(argv[9]) : 4, argc > 10 ? atoi(argv[10]) : 8};
  const bool strict = argc > 11 && atoi(argv[11]) != 0;
  ma_mesh_storage *mh = nullptr;
  auto t0 = std::chrono::steady_clock::now();
  if (ma_mesh_generate(&opt, rank, nranks, &mh)) return printf("mesh: %s\n", ma_last_error()), 1;
  ma::HostLayout A, B;
  if (ma::build_layout(*ma_mesh_view(mh), td, strict, A)) return printf("layout: %s\n", ma_last_error()), 1;
  auto t1 = std::chrono::steady_clock::now();
  if (ma::build_layout_structured(opt, rank, nranks, td, strict, false, B, nullptr))
    return printf("structured layout: %s\n", ma_last_error()), 1;
  auto t2 = std::chrono::steady_clock::now();
  ma::HostLayout C;
  ma::StructuredGrid grid;
  if (ma::build_layout_structured(opt, rank, nranks, td, strict, true, C, &grid))
    return printf("structured layout (deferred geometry): %s\n", ma_last_error()), 1;
  auto t3 = std::chrono::steady_clock::now();
  auto sec = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
  printf("cells %d (+%d ghosts) tiles %d: mesh + layout %.2f s, structured %.2f s, structured without geometry %.2f s\n",
         A.n_owned, A.n_ghost, A.n_tiles, sec(t0, t1), sec(t1, t2), sec(t2, t3));
  long bad = 0;
  {
    const ma::HostLayout &X = A, &Y = B;
    long d = 0;
    d += X.n_owned != Y.n_owned || X.n_ghost != Y.n_ghost || X.stride != Y.stride || X.n_tiles != Y.n_tiles ||
         X.n_interior_tiles != Y.n_interior_tiles || X.n_tile_faces != Y.n_tile_faces ||
         X.n_tile_faces_real != Y.n_tile_faces_real || X.max_tile_faces != Y.max_tile_faces ||
         X.max_tile_cells_real != Y.max_tile_cells_real || X.max_tile_halo != Y.max_tile_halo ||
         X.max_tile_local != Y.max_tile_local || X.halo_stride != Y.halo_stride || X.slot_stride != Y.slot_stride;
    if (d) printf("  DIFFERENT: scalar fields\n");
    d += diff("new2old", X.new2old, Y.new2old) + diff("old2new", X.old2new, Y.old2new);
    d += X.tiles.size() != Y.tiles.size() ||
         (X.tiles.size() && memcmp(X.tiles.data(), Y.tiles.data(), X.tiles.size() * sizeof(ma::TileInfo)) != 0);
    d += diff("cell_xyz", X.cell_xyz, Y.cell_xyz) + diff("cell_vol", X.cell_vol, Y.cell_vol);
    d += diff("slot_face", X.slot_face, Y.slot_face) + diff("slot_nbr", X.slot_nbr, Y.slot_nbr);
    d += diff("face_geom", X.face_geom, Y.face_geom);
    d += diff("face_left", X.face_left, Y.face_left) + diff("face_right", X.face_right, Y.face_right);
    d += diff("face_lr", X.face_lr, Y.face_lr) + diff("tile_halo", X.tile_halo, Y.tile_halo);
    d += diff("send_ids", X.send_ids, Y.send_ids) + diff("recv_ids", X.recv_ids, Y.recv_ids);
    d += diff("peer_rank", X.peer_rank, Y.peer_rank) + diff("peer_send_count", X.peer_send_count, Y.peer_send_count) +
         diff("peer_recv_count", X.peer_recv_count, Y.peer_recv_count);
    if (d && opt.angle != 0.0) {
      // a sheared mesh: the generic builder bins cells by centroid spacing and cuts different (equally valid) tiles
      // than the (i, j, k) bricks; results do not depend on the tiling (tests/test_gpu_parity.py)
      printf("  (expected for a sheared mesh: the array path bins by centroid, the structured path by (i, j, k))\n");
      d = 0;
    }
    bad += d;
  }
  // the deferred-geometry layout: same topology, and its face codes re-evaluated on the host give the same geometry
  bad += diff("deferred: face_lr", B.face_lr, C.face_lr) + diff("deferred: slot_face", B.slot_face, C.slot_face) +
         diff("deferred: tile_halo", B.tile_halo, C.tile_halo) + diff("deferred: new2old", B.new2old, C.new2old);
  long gbad = 0;
  const ma::GridGen &g = grid.gen;
  const long ly = g.b.n[1] + 2, lz = g.b.n[2] + 2;
  const size_t NF = (size_t)B.n_tile_faces;
  for (int k = 0; k < B.n_tiles; ++k) {
    const ma::TileInfo &T = B.tiles[k];
    const size_t fcp = (size_t)((T.face_count + 15) / 16 * 16);
    for (int e = 0; e < T.face_count; ++e) {
      const size_t j = (size_t)T.face_start + e;
      const uint32_t code = C.face_code[j];
      const long lat = code >> 3;
      const int f = code & 7;
      const int ci = (int)(lat / (ly * lz)) - 1, cj = (int)(lat / lz % ly) - 1, ck = (int)(lat % lz) - 1;
      double x[3], n[3], t[3], b[3];
      g.face_geometry(ci, cj, ck, f, x, n, t, b);
      for (int d = 0; d < 3; ++d) {
        if (strict) {
          gbad += B.face_geom[(0 + d) * NF + j] != n[d] || B.face_geom[(3 + d) * NF + j] != t[d] ||
                  B.face_geom[(6 + d) * NF + j] != b[d] || B.face_geom[(9 + d) * NF + j] != x[d];
        } else {
          const size_t base = (size_t)6 * T.face_start + e;
          gbad += B.face_geom[base + (0 + d) * fcp] != n[d] || B.face_geom[base + (3 + d) * fcp] != x[d];
        }
      }
    }
  }
  for (long c = 0; c < (long)B.n_owned + B.n_ghost; ++c) {
    int i, j, k2;
    double xyz[3], vol;
    g.cell_ijk(C.new2old[c], i, j, k2);
    g.cell_geometry(i, j, k2, xyz, &vol);
    for (int d = 0; d < 3; ++d) gbad += B.cell_xyz[(size_t)d * B.stride + c] != xyz[d];
    gbad += B.cell_vol[c] != vol;
  }
  if (gbad) printf("  DIFFERENT: %ld geometry values re-evaluated from the face codes / new2old\n", gbad);
  bad += gbad;
  printf("differences: %ld\n", bad);
  ma_mesh_free(mh);
  return bad != 0;
}
