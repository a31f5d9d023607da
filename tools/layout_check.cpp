// Host-side check of the tile layout (no GPU): invariants of the tile-packed face lists and slot maps, and the
// shared-memory wavefront count the staged flux kernel would see (64-bit loads, half-warp by half-warp, one wavefront
// per distinct double-word modulo 16 ... the model of DESIGN.md §3).
//   g++ -O2 -std=c++17 -fopenmp -Iinclude -Iminiaero_b200/csrc tools/layout_check.cpp miniaero_b200/build/{layout,host_mesh,host_common}.o -o /tmp/layout_check
//   /tmp/layout_check NX NY NZ [tx ty tz] [threads]
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "kernels.h"
#include "layout.h"
#include "miniaero_b200.h"

static int wavefronts16(const int *pos, int n) {  // wavefronts of one half-warp: max multiplicity over 16 residues
  int cnt[16] = {0};
  int mx = 0;
  for (int i = 0; i < n; ++i)
    if (pos[i] >= 0) mx = std::max(mx, ++cnt[pos[i] & 15]);
  return mx;
}

int main(int argc, char **argv) {
  ma_options opt;
  ma_options_default(&opt);
  opt.problem_type = 0;
  opt.lx = 0.3048, opt.ly = 1.0, opt.lz = 1.0;
  opt.nx = argc > 1 ? atoi(argv[1]) : 64, opt.ny = argc > 2 ? atoi(argv[2]) : 32, opt.nz = argc > 3 ? atoi(argv[3]) : 32;
  int td[3] = {argc > 4 ? atoi(argv[4]) : 4, argc > 5 ? atoi(argv[5]) : 4, argc > 6 ? atoi(argv[6]) : 8};
  const int threads = argc > 7 ? atoi(argv[7]) : 128;
  ma_mesh_storage *mh = nullptr;
  if (ma_mesh_generate(&opt, 0, 1, &mh)) return printf("mesh: %s\n", ma_last_error()), 1;
  const ma_mesh *mesh = ma_mesh_view(mh);
  // MINIAERO_CHECK_SHUFFLE=seed: hand the builder the internal faces in a random order, as the reference's mesh
  // generator does (Parallel3DMesh.h:362 shuffles them, unseeded): the layout must not depend on it
  ma_mesh shuffled;
  std::vector<double> sx, sn, st, sb;
  std::vector<int> sc, sf;
  const char *corrupt = getenv("MINIAERO_CHECK_CORRUPT");  // conn | slot | missing | duplicate: a malformed mesh must be refused
  if (corrupt && !getenv("MINIAERO_CHECK_SHUFFLE")) setenv("MINIAERO_CHECK_SHUFFLE", "0", 1);  // work on the copies below
  if (const char *seed = getenv("MINIAERO_CHECK_SHUFFLE")) {
    const ma_faces &F = mesh->internal_faces;
    std::vector<int> perm(F.nfaces);
    for (int i = 0; i < F.nfaces; ++i) perm[i] = i;
    unsigned long long r = 88172645463325252ULL + (unsigned long long)atoll(seed);
    for (int i = F.nfaces - 1; i > 0; --i) {  // Fisher-Yates on xorshift64
      r ^= r << 13, r ^= r >> 7, r ^= r << 17;
      std::swap(perm[i], perm[r % (unsigned long long)(i + 1)]);
    }
    sx.resize(3 * (size_t)F.nfaces), sn.resize(sx.size()), st.resize(sx.size()), sb.resize(sx.size());
    sc.resize(2 * (size_t)F.nfaces), sf.resize(sc.size());
    for (int i = 0; i < F.nfaces; ++i) {
      const size_t j = (size_t)perm[i];
      for (int d = 0; d < 3; ++d) {
        sx[3 * (size_t)i + d] = F.coordinates[3 * j + d], sn[3 * (size_t)i + d] = F.face_normal[3 * j + d];
        st[3 * (size_t)i + d] = F.face_tangent[3 * j + d], sb[3 * (size_t)i + d] = F.face_binormal[3 * j + d];
      }
      for (int d = 0; d < 2; ++d)
        sc[2 * (size_t)i + d] = F.face_cell_conn[2 * j + d], sf[2 * (size_t)i + d] = F.cell_flux_index[2 * j + d];
    }
    shuffled = *mesh;
    shuffled.internal_faces.coordinates = sx.data(), shuffled.internal_faces.face_normal = sn.data();
    shuffled.internal_faces.face_tangent = st.data(), shuffled.internal_faces.face_binormal = sb.data();
    shuffled.internal_faces.face_cell_conn = sc.data(), shuffled.internal_faces.cell_flux_index = sf.data();
    if (corrupt && F.nfaces > 2) {
      const std::string how = corrupt;
      const size_t f = (size_t)F.nfaces / 2;
      if (how == "conn") sc[2 * f + 1] = mesh->num_owned_cells + mesh->num_ghosts + 5;  // a cell that does not exist
      if (how == "slot") sf[2 * f] = 6;                                                 // hex cells have slots 0..5
      if (how == "missing") shuffled.internal_faces.nfaces = F.nfaces - 1;             // two (cell, slot) pairs lose their face
      if (how == "duplicate") {                                                         // one face listed twice, another not at all
        for (int d = 0; d < 2; ++d) sc[2 * f + d] = sc[2 * (f - 1) + d], sf[2 * f + d] = sf[2 * (f - 1) + d];
      }
    }
    mesh = &shuffled;
  }
  ma::HostLayout L;
  const auto t0 = std::chrono::steady_clock::now();
  const bool share = getenv("MINIAERO_CHECK_SHARE") != nullptr;  // shared cut faces (layout.h)
  if (ma::build_layout(*mesh, td, false, L, share)) return printf("layout: %s\n", ma_last_error()), 1;
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("cells %d tiles %d tile faces %ld layout %.2f s\n", L.n_owned, L.n_tiles, L.n_tile_faces_real, sec);
  // ---- invariants
  long bad = 0;
  std::vector<int> seen((size_t)mesh->internal_faces.nfaces, 0);
  for (int k = 0; k < L.n_tiles; ++k) {
    const ma::TileInfo &T = L.tiles[k];
    const int shift = T.cell_start & 1, hb = ((shift + T.cell_count + 1) & ~1);
    for (int e = 0; e < T.face_count; ++e) {
      const size_t j = (size_t)T.face_start + e;
      const int l = L.face_left[j], r = L.face_right[j];
      const unsigned lr = L.face_lr[j];
      const int pl = lr & 0xffff, pr = lr >> 16;
      const bool cut = e >= T.cut_start;
      auto pos_of = [&](int c) { return (c >= T.cell_start && c < T.cell_start + T.cell_count) ? shift + c - T.cell_start : -1; };
      if (r < 0) {  // boundary
        if (cut || pos_of(l) != pl || pr != 0xFFFF - (-1 - r)) ++bad;
      } else if (!cut) {
        if (pos_of(l) != pl || pos_of(r) != pr) ++bad;
      } else {
        const int h = hb + (e - T.cut_start);
        const int out = L.tile_halo[(size_t)T.halo_start + (e - T.cut_start)];
        if (pos_of(l) >= 0) {
          if (pos_of(l) != pl || pr != h || out != r || pos_of(r) >= 0) ++bad;
        } else {
          if (pos_of(r) != pr || pl != h || out != l) ++bad;
        }
      }
    }
    // slot maps: every (cell, slot) points at a face of this tile that has the cell on the stated side
    for (int c = T.cell_start; c < T.cell_start + T.cell_count; ++c)
      for (int s = 0; s < 6; ++s) {
        const unsigned sf = L.slot_face[(size_t)s * L.slot_stride + c];
        const int e = sf & 0x3fff, side = (sf >> 15) & 1;
        if (e >= T.face_count) { ++bad; continue; }
        const size_t j = (size_t)T.face_start + e;
        const int cell = side ? L.face_right[j] : L.face_left[j];
        if (cell != c) ++bad;
        if (((sf >> 14) & 1) != (L.face_right[j] < 0)) ++bad;
      }
  }
  // ---- shared cut faces: every face between two owned cells is evaluated exactly once over all tiles (twice only when
  // the two tiles run in the same flux launch), every imported face has exactly one publisher that targets its slot,
  // the evaluated count is even wherever something is imported, tiles are ordered by launch class
  long evals = 0, imports = 0;
  if (L.share_cut_faces) {
    std::vector<int> published((size_t)std::max(1, L.n_import_areas) * 5 * L.import_capacity, 0);
    std::map<std::pair<int, int>, int> count;  // (left, right) cell pair -> evaluations
    if (L.launch_count[0] + L.launch_count[1] + L.launch_count[2] + L.launch_count[3] != L.n_tiles) ++bad;
    for (int k = 0; k < L.n_tiles; ++k) {
      const ma::TileInfo &T = L.tiles[k];
      const int nimp = T.face_count - T.n_eval;
      if ((T.imp_area >= 0) != (nimp > 0)) ++bad;
      if (nimp > 0 && (T.n_eval & 1)) ++bad;
      if (nimp > L.import_capacity) ++bad;
      std::vector<char> referenced((size_t)T.face_count, 0);
      for (int c = T.cell_start; c < T.cell_start + T.cell_count; ++c)
        for (int s = 0; s < 6; ++s) referenced[L.slot_face[(size_t)s * L.slot_stride + c] & 0x3fff] = 1;
      int dummies = 0;
      for (int e = 0; e < T.face_count; ++e) {
        const size_t j = (size_t)T.face_start + e;
        if (!referenced[e]) { ++dummies; continue; }   // the padding duplicate of an evaluated face
        if (L.face_right[j] < 0) continue;
        if (e < T.n_eval) {
          ++count[{L.face_left[j], L.face_right[j]}];
          ++evals;
        } else {
          ++imports;
        }
        if (e >= T.cut_start && e < T.n_eval) {
          const int pub = L.tile_pub[(size_t)T.halo_start + (e - T.cut_start)];
          if (pub >= 0) {
            if (pub >= (int)published.size()) ++bad; else ++published[pub];
          }
        }
      }
      if (dummies > 1 || (dummies == 1 && nimp == 0)) ++bad;
    }
    for (int k = 0; k < L.n_tiles; ++k) {  // every import slot has exactly one publisher
      const ma::TileInfo &T = L.tiles[k];
      for (int j = 0; j < T.face_count - T.n_eval; ++j)
        if (published[(size_t)T.imp_area * 5 * L.import_capacity + j] != 1) ++bad;
    }
    long twice = 0;
    for (auto &kv : count) {
      if (kv.second < 1 || kv.second > 2) ++bad;
      twice += kv.second == 2;
    }
    if ((long)count.size() != mesh->internal_faces.nfaces) ++bad;   // single domain: every internal face, no ghosts
    printf("shared cut faces: %ld evaluations of %d internal faces (%ld evaluated twice), %ld imports, launches %d %d %d %d\n",
           evals, mesh->internal_faces.nfaces, twice, imports, L.launch_count[0], L.launch_count[1], L.launch_count[2],
           L.launch_count[3]);
  }
  // the gradient sweep's walk through a tile group (interleaved_tile): a permutation for every split, pairs first
  for (int ntiles : {0, 1, 2, 7, 64, L.launch_count[0] + L.launch_count[1]})
    for (int n_first : {0, 1, ntiles / 2, (ntiles + 1) / 2, ntiles - 1, ntiles, L.launch_count[0]}) {
      if (n_first < 0 || n_first > ntiles) continue;
      std::vector<char> hit((size_t)ntiles, 0);
      for (int b = 0; b < ntiles; ++b) {
        const int t = ma::interleaved_tile(b, n_first, ntiles);
        if (t < 0 || t >= ntiles || hit[t]) { ++bad; break; }
        hit[t] = 1;
        const int paired = std::min(n_first, ntiles - n_first);
        if (b < 2 * paired && t != ((b & 1) ? n_first + b / 2 : b / 2)) ++bad;
      }
    }
  printf("invariant violations: %ld\n", bad);
  {  // every array of the layout in one number: the builder's result must not depend on the number of host threads
    unsigned long long h = 1469598103934665603ULL;
    auto mix = [&](const void *p, size_t bytes) {
      const unsigned char *c = static_cast<const unsigned char *>(p);
      for (size_t i = 0; i < bytes; ++i) h = (h ^ c[i]) * 1099511628211ULL;
    };
    auto vec = [&](const auto &v) {
      if (!v.empty()) mix(v.data(), v.size() * sizeof(v[0]));
    };
    vec(L.new2old), vec(L.old2new), vec(L.tiles), vec(L.cell_xyz), vec(L.cell_vol), vec(L.slot_face), vec(L.slot_nbr);
    vec(L.face_geom), vec(L.face_left), vec(L.face_right), vec(L.face_lr), vec(L.tile_halo), vec(L.tile_pub);
    printf("layout checksum: %016llx\n", h);
  }
  // ---- wavefront model of the staged flux kernel
  long ideal1 = 0, wf1 = 0, ideal2 = 0, wf2 = 0, paths = 0, warps = 0;
  for (int k = 0; k < L.n_tiles; ++k) {
    const ma::TileInfo &T = L.tiles[k];
    const int nh = T.n_eval - T.cut_start, nf = T.n_eval;  // the faces this tile's CTA evaluates
    const int shift = T.cell_start & 1, hb = ((shift + T.cell_count + 1) & ~1);
    for (int w0 = 0; w0 < nf; w0 += 16) {  // work items of one half-warp (threads is a multiple of 16)
      int p[4][16];
      for (int st = 0; st < 4; ++st)
        for (int i = 0; i < 16; ++i) p[st][i] = -1;
      for (int i = 0; i < 16 && w0 + i < nf; ++i) {
        const int w = w0 + i;
        const int e = w < nh ? T.cut_start + w : w - nh;
        const size_t j = (size_t)T.face_start + e;
        const unsigned lr = L.face_lr[j];
        const int pl = lr & 0xffff, pr = lr >> 16;
        if (w < nh) {
          if (pr >= hb) p[2][i] = pl; else p[3][i] = pr;
        } else if (pr >= 0xFFF0) {
          p[2][i] = pl;
        } else {
          p[0][i] = pl, p[1][i] = pr;
        }
      }
      for (int st = 0; st < 4; ++st) {
        const int m = wavefronts16(p[st], 16);
        wf1 += m;
        ideal1 += m > 0;
      }
    }
    for (int w0 = 0; w0 < nf; w0 += 32) {  // code paths per warp: cut-left, cut-right, interior, boundary
      int kinds = 0;
      for (int i = 0; i < 32 && w0 + i < nf; ++i) {
        const int w = w0 + i;
        const int e = w < nh ? T.cut_start + w : w - nh;
        const unsigned lr = L.face_lr[(size_t)T.face_start + e];
        const int pr = lr >> 16;
        kinds |= w < nh ? (pr >= hb ? 1 : 2) : (pr >= 0xFFF0 ? 8 : 4);
      }
      paths += __builtin_popcount(kinds);
      ++warps;
    }
    for (int c0 = 0; c0 < T.cell_count; c0 += 16)
      for (int s = 0; s < 6; ++s) {
        int p[16];
        for (int i = 0; i < 16; ++i)
          p[i] = c0 + i < T.cell_count ? (int)(L.slot_face[(size_t)s * L.slot_stride + T.cell_start + c0 + i] & 0x3fff) : -1;
        wf2 += wavefronts16(p, 16);
        ideal2 += 1;
      }
  }
  // gradient kernel: thread per own cell; per slot it reads 6 geometry values of face e(s, c) and 5 primitives of
  // the neighbour at staged position nb(s, c)
  long gi = 0, gw_e = 0, gw_n = 0;
  for (int k = 0; k < L.n_tiles; ++k) {
    const ma::TileInfo &T = L.tiles[k];
    for (int c0 = 0; c0 < T.cell_count; c0 += 16)
      for (int s = 0; s < 6; ++s) {
        int pe[16], pn[16];
        for (int i = 0; i < 16; ++i) {
          const bool in = c0 + i < T.cell_count;
          const size_t idx = (size_t)s * L.slot_stride + T.cell_start + c0 + i;
          pe[i] = in ? (int)(L.slot_face[idx] & 0x3fff) : -1;
          pn[i] = in && L.slot_nbr[idx] != 0xFFFF ? (int)L.slot_nbr[idx] : -1;
        }
        gw_e += wavefronts16(pe, 16);
        gw_n += std::max(1, wavefronts16(pn, 16));
        ++gi;
      }
  }
  printf("gradient kernel: geometry-read wavefronts / ideal %.3f, neighbour-read %.3f, weighted (6:5) %.3f\n",
         (double)gw_e / gi, (double)gw_n / gi, (6.0 * gw_e + 5.0 * gw_n) / (11.0 * gi));
  (void)threads;
  printf("code paths per warp of 32 work items: %.3f\n", (double)paths / warps);
  printf("phase 1 record-read wavefronts / ideal: %.3f   phase 2 flux-read wavefronts / ideal: %.3f\n",
         (double)wf1 / ideal1, (double)wf2 / ideal2);
  printf("weighted (28 reads per record stream, 5 per flux read): %.3f\n",
         (28.0 * wf1 + 5.0 * wf2) / (28.0 * ideal1 + 5.0 * ideal2));
  ma_mesh_free(mh);
  return bad != 0;
}
