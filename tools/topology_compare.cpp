// Host-side check (no GPU) of the device topology builder's logic: build_topology_plan() + the stamping functions of
// topology_stamp.h (the very functions the CUDA kernels of topology_kernels.cu run, executed here in plain loops)
// against build_layout_structured(), the host builder.  Every array must be identical, bit for bit.
//   topology_compare NX NY NZ PROBLEM_TYPE RANK NRANKS [tx ty tz] [share]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "layout.h"
#include "miniaero_b200.h"
#include "topology_stamp.h"

template <class T, class A, class B>
static long diff(const char *name, const std::vector<T, A> &a, const std::vector<T, B> &b) {
  long bad = a.size() != b.size();
  if (!bad) bad = a.empty() ? 0 : memcmp(a.data(), b.data(), a.size() * sizeof(T)) != 0;
  if (bad) {
    size_t first = 0;
    while (first < a.size() && first < b.size() && a[first] == b[first]) ++first;
    printf("  DIFFERENT: %s (%zu vs %zu entries, first difference at %zu)\n", name, a.size(), b.size(), first);
  }
  return bad;
}

int main(int argc, char **argv) {
  if (argc < 7) return printf("usage: topology_compare NX NY NZ PTYPE RANK NRANKS [tx ty tz] [share]\n"), 2;
  ma_options opt;
  ma_options_default(&opt);
  opt.nx = atoi(argv[1]), opt.ny = atoi(argv[2]), opt.nz = atoi(argv[3]);
  opt.problem_type = atoi(argv[4]);
  const int rank = atoi(argv[5]), nranks = atoi(argv[6]);
  opt.lx = opt.problem_type == 0 ? 0.3048 : 2.0, opt.ly = opt.problem_type == 1 ? 0.002 : 1.3, opt.lz = 0.9;
  int td[3] = {argc > 9 ? atoi(argv[7]) : 4, argc > 9 ? atoi(argv[8]) : 4, argc > 9 ? atoi(argv[9]) : 8};
  const bool share = argc > 10 && atoi(argv[10]) != 0;
  ma::HostLayout A, B;
  ma::StructuredGrid ga, gb;
  auto t0 = std::chrono::steady_clock::now();
  if (ma::build_layout_structured(opt, rank, nranks, td, false, true, A, &ga, share)) return printf("host layout: %s\n", ma_last_error()), 1;
  auto t1 = std::chrono::steady_clock::now();
  ma::TopoPlan P;
  if (ma::build_topology_plan(opt, rank, nranks, td, share, B, &gb, P)) return printf("plan: %s\n", ma_last_error()), 1;
  auto t2 = std::chrono::steady_clock::now();

  // flatten the plan the way solver.cu uploads it
  std::vector<ma::TileInfoDev> tiles(B.tiles.size());
  for (size_t i = 0; i < tiles.size(); ++i)
    tiles[i] = {B.tiles[i].cell_start, B.tiles[i].cell_count, B.tiles[i].face_start, B.tiles[i].face_count,
                B.tiles[i].cut_start, B.tiles[i].n_eval, B.tiles[i].imp_area, 0};
  std::vector<int> pat_ext, pat_dummy, pat_cell_off, pat_rank_off, pat_face_off;
  std::vector<uint32_t> cell_abc;
  std::vector<uint16_t> rank_of, face_lc;
  std::vector<uint8_t> face_slot;
  for (const ma::TopoPattern &p : P.patterns) {
    for (int d = 0; d < 3; ++d) pat_ext.push_back(p.ext[d]);
    pat_dummy.push_back(p.dummy_face);
    pat_cell_off.push_back((int)cell_abc.size());
    pat_rank_off.push_back((int)rank_of.size());
    pat_face_off.push_back((int)face_lc.size());
    cell_abc.insert(cell_abc.end(), p.cell_abc.begin(), p.cell_abc.end());
    rank_of.insert(rank_of.end(), p.rank_of.begin(), p.rank_of.end());
    face_lc.insert(face_lc.end(), p.face_lc.begin(), p.face_lc.end());
    face_slot.insert(face_slot.end(), p.face_slot.begin(), p.face_slot.end());
  }
  const long n_cells = (long)B.n_owned + B.n_ghost;
  std::vector<int> new2old(n_cells), old2new(n_cells), tile_halo((size_t)B.n_tiles * B.halo_stride, -1), tile_pub;
  if (share) tile_pub.assign((size_t)B.n_tiles * B.halo_stride, -1);
  for (long c = B.n_owned; c < n_cells; ++c) new2old[c] = old2new[c] = (int)c;
  std::vector<uint16_t> slot_face((size_t)6 * B.slot_stride, 0), slot_nbr((size_t)6 * B.slot_stride, 0xFFFF);
  std::vector<uint32_t> face_lr((size_t)B.n_tile_faces, 0), face_code((size_t)B.n_tile_faces, 0);
  ma::TopoView t;
  t.g = gb.gen;
  t.n_owned = B.n_owned, t.n_tiles = B.n_tiles, t.slot_stride = B.slot_stride, t.halo_stride = B.halo_stride;
  t.import_capacity = B.import_capacity;
  for (int f = 0; f < 6; ++f) t.bc_of_face[f] = P.bc_of_face[f];
  t.tiles = tiles.data(), t.tile_pattern = P.tile_pattern.data(), t.tile_origin = P.tile_origin.data();
  t.tile_nb = P.tile_nb.data(), t.tile_launch = P.tile_launch.data();
  t.pat_ext = pat_ext.data(), t.pat_dummy = pat_dummy.data(), t.pat_cell_off = pat_cell_off.data();
  t.pat_rank_off = pat_rank_off.data(), t.pat_face_off = pat_face_off.data();
  t.cell_abc = cell_abc.data(), t.rank_of = rank_of.data(), t.face_lc = face_lc.data(), t.face_slot = face_slot.data();
  t.new2old = new2old.data(), t.old2new = old2new.data(), t.slot_face = slot_face.data(), t.slot_nbr = slot_nbr.data();
  t.face_lr = face_lr.data(), t.face_code = face_code.data(), t.tile_halo = tile_halo.data();
  t.tile_pub = share ? tile_pub.data() : nullptr;
  for (int k = 0; k < B.n_tiles; ++k)
    for (int lc = 0; lc < tiles[k].cell_count; ++lc) ma::topo_stamp_cell(t, k, lc);
  for (int k = 0; k < B.n_tiles; ++k)
    for (int e = 0; e < tiles[k].face_count; ++e) ma::topo_stamp_face(t, k, e);
  if (share)
    for (int k = 0; k < B.n_tiles; ++k)
      for (int q = 0; q < tiles[k].n_eval - tiles[k].cut_start; ++q) ma::topo_stamp_pub(t, k, q);
  auto t3 = std::chrono::steady_clock::now();

  long bad = 0;
#define SCALAR(f) if (A.f != B.f) { printf("  DIFFERENT: %s (%ld vs %ld)\n", #f, (long)A.f, (long)B.f); ++bad; }
  SCALAR(n_owned) SCALAR(n_ghost) SCALAR(stride) SCALAR(n_tiles) SCALAR(n_interior_tiles) SCALAR(n_tile_faces)
  SCALAR(n_tile_faces_real) SCALAR(max_tile_cells) SCALAR(max_tile_cells_real) SCALAR(max_tile_faces) SCALAR(max_tile_local)
  SCALAR(max_tile_halo) SCALAR(halo_stride) SCALAR(slot_stride) SCALAR(import_capacity) SCALAR(n_import_areas)
  SCALAR(share_cut_faces) SCALAR(geom_components)
  for (int i = 0; i < 4; ++i) SCALAR(launch_count[i])
#undef SCALAR
  for (int k = 0; k < std::min(A.n_tiles, B.n_tiles); ++k) {
    const ma::TileInfo &x = A.tiles[k], &y = B.tiles[k];
    if (x.cell_start != y.cell_start || x.cell_count != y.cell_count || x.face_start != y.face_start ||
        x.face_count != y.face_count || x.cut_start != y.cut_start || x.halo_start != y.halo_start || x.n_eval != y.n_eval ||
        x.imp_area != y.imp_area) {
      printf("  DIFFERENT: tile %d\n", k);
      ++bad;
      break;
    }
  }
  bad += diff("new2old", A.new2old, new2old);
  bad += diff("old2new", A.old2new, old2new);
  bad += diff("slot_face", A.slot_face, slot_face);
  bad += diff("slot_nbr", A.slot_nbr, slot_nbr);
  bad += diff("face_lr", A.face_lr, face_lr);
  bad += diff("face_code", A.face_code, face_code);
  bad += diff("tile_halo", A.tile_halo, tile_halo);
  bad += diff("tile_pub", A.tile_pub, tile_pub);
  bad += diff("send_ids", A.send_ids, B.send_ids);
  bad += diff("recv_ids", A.recv_ids, B.recv_ids);
  bad += diff("peer_rank", A.peer_rank, B.peer_rank);
  bad += diff("peer_send_count", A.peer_send_count, B.peer_send_count);
  bad += diff("peer_recv_count", A.peer_recv_count, B.peer_recv_count);
  auto sec = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
  printf("cells %d (+%d ghosts) tiles %d patterns %zu: host builder %.3f s, plan %.3f s, stamping (one thread) %.3f s\n",
         A.n_owned, A.n_ghost, A.n_tiles, P.patterns.size(), sec(t0, t1), sec(t1, t2), sec(t2, t3));
  printf("differences: %ld\n", bad);
  return bad != 0;
}
